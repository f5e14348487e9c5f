// tools/emu: empty stand-in so that <nccl.h> parses under the host compiler (test infrastructure)
