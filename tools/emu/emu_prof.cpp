// tools/emu: C entry points of the emulator's section profile (test infrastructure)
#include "cuda_runtime.h"
extern "C" void zrt_emu_prof_reset() { zrt_emu::g_prof = zrt_emu::Prof{}; }
extern "C" void zrt_emu_prof_get(unsigned long long *calls, unsigned long long *lanes) {
    for (int i = 0; i < 64; i++) { calls[i] = zrt_emu::g_prof.calls[i]; lanes[i] = zrt_emu::g_prof.lanes[i]; }
}
