"""zraytrace_b200 — host-side mirror of the zraytrace render boundary over libzrt.so (sm_100a CUDA).

    from zraytrace_b200 import host, lib, make_params
    scene = host.HostScene(host.SCENE_THREE_BALLS)            # scenes.zig threeBalls
    with lib.Scene(scene, device=0) as dev:                   # flatten + upload (raytrace.zig:150)
        image, counters, timing = dev.render(scene.camera, make_params(1000, 1000, 1000, 30))

There is no CPU fallback: every compute call goes through the CUDA library or raises."""
from ._abi import Camera, Counters, Params, Timing, make_params  # noqa: F401
