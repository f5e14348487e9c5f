"""world_size-2 (and 3) `gloo` tests of the multi-GPU host logic on CPU: sample-range partition, raw-sum
reduction and final 1/spp scaling (zraytrace_b200/distributed.py).  The per-rank "renderer" here is the CPU
oracle restricted to the rank's global sample range, standing in for zrt_render_device, which needs a GPU;
everything else is the code path bench.py and render_distributed use."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zraytrace_b200 import _abi as A
from zraytrace_b200 import distributed as D


def test_sample_ranges_partition_exactly():
    for spp in (1, 7, 100, 1000, 1024):
        for world in (1, 2, 3, 4, 8):
            r = [D.sample_range(spp, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == spp
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    p = A.make_params(8, 8, 10, 3)
    q = D.rank_params(p, 1, 4)
    assert (q.sample_begin, q.sample_end) == (2, 5) and q.flags & A.ZRT_FLAG_RAW_SUM and p.flags == 0
    assert D.color_scale(1000) == np.float32(1.0) / np.float32(1000)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import zro_py
    from tests import scenes_py

    sc, cam = scenes_py.three_balls()
    params = A.make_params(24, 24, 10, 30)
    p = D.rank_params(params, rank, world)
    img, cnt, _ = zro_py.render(sc, cam, p, rng=zro_py.RNG_CTR)  # raw sums over this rank's global samples
    accum = torch.from_numpy(img.copy())
    counters = torch.tensor(list(cnt.as_dict().values()), dtype=torch.int64)
    is_dst = D.reduce_and_scale(accum, counters, params.samples_per_pixel)
    if rank == 0:
        assert is_dst
        np.save(out + ".img.npy", accum.numpy())
        np.save(out + ".cnt.npy", counters.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_spp_split_reduce_matches_single_process(world, tmp_path):
    from oracle import zro_py
    from tests import scenes_py

    out = str(tmp_path / "r")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    img = np.load(out + ".img.npy")
    cnt = np.load(out + ".cnt.npy")
    sc, cam = scenes_py.three_balls()
    full, c_full, _ = zro_py.render(sc, cam, A.make_params(24, 24, 10, 30), rng=zro_py.RNG_CTR)
    want = np.array(list(c_full.as_dict().values()), dtype=np.int64)
    want[3] = c_full.pixels_processed  # pixels are counted by the rank that owns sample 0 only
    assert np.array_equal(cnt, want)  # u64 counters are exactly world-size invariant
    np.testing.assert_allclose(img, full, rtol=1e-5, atol=1e-6)  # f32 sums differ only by association
