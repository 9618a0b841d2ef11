"""The exchange step fused with the optimiser (``eavqa_sharded_adamw_step``, csrc/collective.cu): reduce-scatter + AdamW +
all-gather as one kernel per rank over peer pointers.

CPU: the shard arithmetic (host mirror against the C ABI) and, over gloo with two processes, the semantics the kernel
implements -- every rank updating only its shard of the averaged gradient and broadcasting it reproduces all-reduce + AdamW
on every rank.  GPU (one device): the kernel itself, with two simulated ranks whose buffers live on the same GPU and whose
kernels run concurrently on two streams (the barriers inside the kernel are real), against ``eavqa_adamw_step`` and
``torch.optim.AdamW`` -- what Lightning DDP + ``ClipCapExecutor.configure_optimizers`` compute (main.py:133-138,
clipcap_exector.py:79-81).  The NVLS multicast path needs more than one GPU: ``tools/check_sharded_adamw.py`` under torchrun.
"""
import ctypes as C
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eavqa_b200 import lib as L_
from eavqa_b200.parallel import bucket_plan, owned_ranges, shard_range


def test_shard_range_matches_the_c_abi_and_tiles_the_buffer():
    L = L_.load()
    for n in (0, 4, 8, 40, 4096, 41_700_000, 41_700_004):
        for world in (1, 2, 3, 4, 8, 16):
            pos = 0
            for r in range(world):
                b, e = C.c_int64(), C.c_int64()
                L_.check(L.eavqa_sharded_adamw_range(n, r, world, C.byref(b), C.byref(e)))
                assert (b.value, e.value) == shard_range(n, r, world)
                assert b.value == pos and e.value >= b.value and b.value % 4 == 0 and e.value % 4 == 0
                pos = e.value
            assert pos == n
    with pytest.raises(ValueError):
        shard_range(6, 0, 2)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_owned_ranges_tile_the_buffer_bucket_by_bucket():
    """Per-bucket exchange: every rank owns its shard of every bucket; over the ranks these tile the whole flat buffer."""
    n = 4 * 1000
    buckets, rest = bucket_plan(n, [(2400, 3600), (1200, 2400), (400, 1200)])     # completion order; [0, 400) and [3600, n) uncovered
    assert rest == [(0, 400), (3600, 4000)]
    for world in (1, 2, 3, 8):
        cover = []
        for r in range(world):
            mine = owned_ranges(buckets + rest, r, world)
            assert all(b % 4 == 0 and e % 4 == 0 and e > b for b, e in mine)
            cover += mine
        cover.sort()
        assert cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    assert owned_ranges([(0, 8)], 3, 4) == []          # 2 float4 over 4 ranks: the last ranks own nothing


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _adamw_ref(p, g, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=0.01):
    p = p * (1 - lr * wd)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    denom = v.sqrt() / (1 - b2 ** step) ** 0.5 + eps
    return p - (lr / (1 - b1 ** step)) * (m / denom), m, v


def _gloo_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 4 * 37
    torch.manual_seed(0)
    p0 = torch.randn(n, dtype=torch.float64)
    params = p0.clone()
    m, v = torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    ref_p, ref_m, ref_v = p0.clone(), m.clone(), v.clone()
    b, e = shard_range(n, rank, world)
    for step in (1, 2, 3):
        g_local = torch.randn(n, dtype=torch.float64, generator=torch.Generator().manual_seed(100 * step + rank))
        # what the kernel does: sum over ranks of the OWN shard, 1/W, AdamW on the shard, shard stored to every rank
        shards = [torch.zeros(shard_range(n, r, world)[1] - shard_range(n, r, world)[0], dtype=torch.float64) for r in range(world)]
        for r in range(world):
            rb, re_ = shard_range(n, r, world)
            piece = g_local[rb:re_].clone()
            dist.reduce(piece, dst=r)
            if r == rank:
                shards[r] = piece
        new_p, m[b:e], v[b:e] = _adamw_ref(params[b:e], shards[rank] / world, m[b:e], v[b:e], step)
        gathered = [torch.zeros_like(s) for s in shards]
        for r in range(world):
            t = new_p.clone() if r == rank else gathered[r]
            dist.broadcast(t, src=r)
            gathered[r] = t
        params = torch.cat(gathered)
        # the reference: all-reduce mean, AdamW everywhere
        g_all = g_local.clone()
        dist.all_reduce(g_all)
        ref_p, ref_m, ref_v = _adamw_ref(ref_p, g_all / world, ref_m, ref_v, step)
    assert torch.allclose(params, ref_p, rtol=0, atol=1e-12)
    assert torch.allclose(m[b:e], ref_m[b:e], rtol=0, atol=1e-14)
    if rank == 0:
        torch.save(params, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_update_equals_allreduce_then_adamw_gloo(tmp_path):
    out = str(tmp_path / "p.pt")
    mp.spawn(_gloo_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out).shape == (4 * 37,)


# ------------------------------------------------------------------------------------------------ GPU
def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[C.c_void_p(t.data_ptr()) for t in tensors])


@pytest.mark.gpu
@pytest.mark.parametrize("world,n,inkernel,ranges", [(1, 4096, False, None), (1, 4096, True, None), (2, 40_000, True, None),
                                                    (4, 100_004, True, None), (2, 8, True, None),
                                                    (2, 40_000, True, [(30_000, 10_000), (8_000, 22_000), (0, 8_000)]),
                                                    (3, 20_008, True, [(20_000, 8), (4, 19_996), (0, 4)])])
def test_sharded_adamw_simulated_ranks_on_one_device(world, n, inkernel, ranges):
    L = L_.load()
    dev = "cuda"
    torch.manual_seed(3)
    p0 = torch.randn(n, device=dev)
    params = [p0.clone() for _ in range(world)]
    grads = [torch.zeros(n, device=dev) for _ in range(world)]
    flags = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(world)]
    m = [torch.zeros(n, device=dev) for _ in range(world)]
    v = [torch.zeros(n, device=dev) for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    ref_p, ref_m, ref_v = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    tp = torch.nn.Parameter(p0.clone())
    topt = torch.optim.AdamW([tp], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    gp, pp, fp = _ptr_array(grads), _ptr_array(params), _ptr_array(flags)
    for step in (1, 2, 3):
        total = torch.zeros(n, device=dev)
        for r in range(world):
            grads[r].copy_(torch.randn(n, device=dev) * 10.0 ** -step)
            total += grads[r]                                 # rank order, like the peer-load path
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                if ranges is None:
                    L_.check(L.eavqa_sharded_adamw_step(gp, pp, None, None, fp if inkernel else None, step, r, world, m[r].data_ptr(),
                                                        v[r].data_ptr(), 0, n, 0, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, 1.0 / world,
                                                        streams[r].cuda_stream))
                else:       # the buffer exchanged range by range (gradient buckets), a few CTAs each, one barrier generation per call
                    for k, (off, cnt) in enumerate(ranges):
                        L_.check(L.eavqa_sharded_adamw_step(gp, pp, None, None, fp, (step - 1) * len(ranges) + k + 1, r, world,
                                                            m[r].data_ptr(), v[r].data_ptr(), off, cnt, 2, 1e-3, 0.9, 0.999, 1e-8, 0.01,
                                                            step, 1.0 / world, streams[r].cuda_stream))
        torch.cuda.synchronize()
        # the library's own fused AdamW on the summed gradient (bit-exact: same arithmetic, same order)
        L_.check(L.eavqa_adamw_step(ref_p.data_ptr(), total.data_ptr(), ref_m.data_ptr(), ref_v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8,
                                    0.01, step, 1.0 / world, torch.cuda.current_stream().cuda_stream))
        tp.grad = total / world
        topt.step()
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(params[r], ref_p), "rank %d's replica differs after step %d" % (r, step)
            assert int(flags[r][33]) == 0, "a barrier spin timed out"
    # torch.optim.AdamW: what the reference's executor builds
    assert (params[0] - tp.detach()).abs().max().item() <= 2e-6
    for r in range(world):
        own = torch.zeros(n, dtype=torch.bool, device=dev)
        for off, cnt in (ranges if ranges is not None else [(0, n)]):
            b, e = shard_range(cnt, r, world)
            own[off + b:off + e] = True
        assert torch.equal(m[r][own], ref_m[own]) and torch.equal(v[r][own], ref_v[own])
        assert float(m[r][~own].abs().sum()) == 0.0             # only the own shards of the moments are touched


@pytest.mark.gpu
def test_sharded_adamw_rejects_bad_arguments():
    L = L_.load()
    t = torch.zeros(8, device="cuda")
    arr = _ptr_array([t])
    s = torch.cuda.current_stream().cuda_stream
    assert L.eavqa_sharded_adamw_step(arr, arr, None, None, None, 1, 0, 1, t.data_ptr(), t.data_ptr(), 0, 6, 0, 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, 1.0, s) != 0
    assert L.eavqa_sharded_adamw_step(arr, arr, None, None, None, 1, 1, 1, t.data_ptr(), t.data_ptr(), 0, 8, 0, 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, 1.0, s) != 0
    assert L.eavqa_sharded_adamw_step(arr, arr, t.data_ptr(), None, None, 1, 0, 1, t.data_ptr(), t.data_ptr(), 0, 8, 0, 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, 1.0, s) != 0
    assert b"sharded_adamw_step" in L.eavqa_last_error()
