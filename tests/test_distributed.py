"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes, the shard
arithmetic, the packed accumulator layout and the single all-reduce.  The compute itself has no CPU
path (by design), so a stand-in engine writes per-shard accumulators whose global sum is known."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from sdetools_b200 import distributed as sdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_ensemble():
    for n in (0, 1, 7, 1000, 10 ** 10 + 3):
        for world in (1, 2, 3, 8):
            r = [sdist.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(hi >= lo for lo, hi in r)
            assert max(hi - lo for lo, hi in r) <= -(-n // world)


def test_pack_unpack_roundtrip():
    ps = np.arange(10.0).reshape(5, 2, order="F")
    hc = np.array([1, 2, 3, 2 ** 52, 5, 0, 7], dtype=np.uint64)
    buf = sdist.pack_stats(ps, hc)
    assert buf.shape == (17,)
    ps2, hc2 = sdist.unpack_stats(buf, nx=2, nbins=4)
    assert np.array_equal(ps2, ps) and np.array_equal(hc2, hc)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from sdetools_b200 import distributed as sdist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    class FakeEngine:                      # stands in for libsdegpu: accumulators are functions of the path ids
        device = 0
        def torch_device(self):
            return torch.device("cpu")
        def run(self, model, scheme, params, times, x0, *, nx, npaths, projection, seed, path_offset, power_sums,
                center, hist_counts, hist, want_time):
            ids = torch.arange(path_offset, path_offset + npaths, dtype=torch.float64)
            power_sums.copy_(torch.tensor([npaths, ids.sum(), (ids ** 2).sum(), 0.0, 0.0], dtype=torch.float64))
            if hist_counts is not None:
                hist_counts.zero_()
                hist_counts[1:hist[2] + 1] = torch.bincount((ids.long() %% hist[2]), minlength=hist[2])
            return None

    N = 1001
    packed, mine, _ = sdist.run_sharded(FakeEngine(), "ou", "euler", [1, 0, 1], [0.0, 1.0], [0.0], nx=1, npaths_total=N,
                                        hist=(0.0, 1.0, 8), rank=rank, world=world)
    ps, hc = sdist.unpack_stats(packed.numpy(), nx=1, nbins=8)
    ids = np.arange(N, dtype=np.float64)
    assert ps[0, 0] == N and ps[1, 0] == ids.sum() and ps[2, 0] == (ids ** 2).sum(), ps
    assert hc[1:9].tolist() == np.bincount(np.arange(N) %% 8, minlength=8).tolist() and hc.sum() == N
    lo, hi = sdist.shard_range(N, rank, world)
    assert mine == hi - lo
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_two_rank_gloo_allreduce_of_accumulators(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29611", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
