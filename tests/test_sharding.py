"""Multi-rank host logic: partition arithmetic, halo rows and the result gather, on CPU with the gloo
backend (world_size 2 and 3).  The per-rank compute is the oracle here (tests may use it); on the GPU
box the same ShardedStereo runs on GpuCompute (test_gpu_sharding below, 1 GPU, world_size 1)."""
import os
import socket

import numpy as np
import pytest

import oracle
import introtocomputervision_b200 as sb
from introtocomputervision_b200 import sharding, synth


def test_pair_shard_partitions_exactly():
    for n in (0, 1, 5, 8, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [sharding.pair_shard(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_band_shard_and_halo():
    for rows in (1, 7, 270, 2160):
        for world in (1, 2, 4, 8):
            spans = [sharding.band_shard(rows, world, r) for r in range(world)]
            assert spans[0][0] == 0 and max(e for _, e in spans) == rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    # 2160 rows on 8 ranks, 11x11 window: 270-row bands, R+1 = 6 halo rows each side, clamped at the image edge
    assert sharding.band_shard(2160, 8, 3) == (810, 1080)
    assert sharding.band_halo(2160, 810, 1080, 5) == (804, 1086)
    assert sharding.band_halo(2160, 0, 270, 5) == (0, 276)
    assert sharding.band_halo(2160, 1890, 2160, 5) == (1884, 2160)
    with pytest.raises(ValueError):
        sharding.band_halo(10, 5, 3, 1)


class OracleCompute:
    """Stand-in for GpuCompute: runs the oracle on the slab as if it were the whole image and keeps the band.
    With the R+1 halo this equals the full-image result on the band's rows."""
    device = "cpu"

    def band(self, cost, left_slab, right_slab, rows, cols, r0, r1, h0, h1, R, dmin, dmax, dtype):
        import torch
        fn = oracle.ssd_fast if cost == sb.COST_SSD else oracle.ncorr_fast
        d = fn(left_slab.astype(np.float32), right_slab.astype(np.float32), R, dmin, dmax)
        return torch.from_numpy(d[r0 - h0:r1 - h0].astype(np.int32)).to(dtype)

    def pair_band(self, cost, left_slab, right_slab, rows, cols, r0, r1, h0, h1, R, rng, dtype):
        import torch
        return torch.stack([self.band(cost, left_slab, right_slab, rows, cols, r0, r1, h0, h1, R, -rng, 0, dtype),
                            self.band(cost, right_slab, left_slab, rows, cols, r0, r1, h0, h1, R, 0, rng, dtype)])

    def pair_batch(self, cost, lefts, rights, R, rng, dtype):
        import torch
        fn = oracle.ssd_fast if cost == sb.COST_SSD else oracle.ncorr_fast
        out = np.zeros((2,) + lefts.shape, np.int32)
        for i in range(lefts.shape[0]):
            a, b = lefts[i].astype(np.float32), rights[i].astype(np.float32)
            out[0, i], out[1, i] = fn(a, b, R, -rng, 0), fn(b, a, R, 0, rng)
        return torch.from_numpy(out).to(dtype)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle.set_num_threads(1)
        sh = sharding.ShardedStereo(OracleCompute())
        ok = True
        # row bands: 45 rows do not divide by 2 or 3 -> the last band is shorter
        L, Rt, _ = synth.make_pair(45, 96, 12, 77)
        for cost, fn in ((sb.COST_SSD, oracle.ssd_fast), (sb.COST_NCORR, oracle.ncorr_fast)):
            for (a, b, dmin, dmax) in ((L, Rt, -11, 0), (Rt, L, 0, 11)):
                full = sh.disparity_bands(cost, a, b, 3, dmin, dmax, dtype=torch.int16).numpy()
                ref = fn(a.astype(np.float32), b.astype(np.float32), 3, dmin, dmax)
                ok &= bool(np.array_equal(full, ref.astype(np.int16)))
        # both maps of a pair, row-band sharded
        pl, pr = sh.disparity_pair_bands(sb.COST_SSD, L, Rt, 3, 11, dtype=torch.int16)
        ok &= bool(np.array_equal(pl.numpy(), oracle.ssd_fast(L.astype(np.float32), Rt.astype(np.float32), 3, -11, 0).astype(np.int16)))
        ok &= bool(np.array_equal(pr.numpy(), oracle.ssd_fast(Rt.astype(np.float32), L.astype(np.float32), 3, 0, 11).astype(np.int16)))
        # pair batch: 5 pairs over the ranks
        Ls = np.stack([synth.make_pair(20, 64, 8, 300 + i)[0] for i in range(5)])
        Rs = np.stack([synth.make_pair(20, 64, 8, 300 + i)[1] for i in range(5)])
        dl, dr = sh.disparity_pair_batch(sb.COST_SSD, Ls, Rs, 2, 7, dtype=torch.int8)
        for i in range(5):
            a, b = Ls[i].astype(np.float32), Rs[i].astype(np.float32)
            ok &= bool(np.array_equal(dl[i].numpy(), oracle.narrow_i8(oracle.ssd_fast(a, b, 2, -7, 0))))
            ok &= bool(np.array_equal(dr[i].numpy(), oracle.narrow_i8(oracle.ssd_fast(b, a, 2, 0, 7))))
        # peer gather set-up is collective: without a device no rank can create a buffer, and EVERY rank gets the
        # same PeerGatherUnavailable instead of one rank raising while the others wait in an exchange
        class NoDeviceCtx:
            handle = None
        try:
            sharding.PeerGather(NoDeviceCtx(), 1024)
            ok = False
        except sharding.PeerGatherUnavailable:
            pass
        dist.barrier()
        q.put((rank, ok))
    except Exception as exc:      # report instead of leaving the parent waiting on the queue
        q.put((rank, f"{type(exc).__name__}: {exc}"))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_bands_and_batches_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in results) == list(range(world))
    assert all(ok is True for _, ok in results), results


@pytest.mark.gpu
def test_gpu_sharding_single_rank_and_slab_api(ctx):
    """GpuCompute through the slab (halo) entry point: every band of a 4-way split, computed from its slab
    only, equals the full-image call."""
    import torch
    comp = sharding.GpuCompute(ctx, torch.device("cuda", 0))
    L, Rt, _ = synth.make_pair(203, 300, 40, 5)
    for cost in (sb.COST_SSD, sb.COST_NCORR):
        for (a, b, dmin, dmax) in ((L, Rt, -39, 0), (Rt, L, 0, 39)):
            full = ctx.disparity(cost, a, b, 5, dmin, dmax, dtype=np.int16)
            for rank in range(4):
                r0, r1 = sharding.band_shard(203, 4, rank)
                h0, h1 = sharding.band_halo(203, r0, r1, 5)
                band = comp.band(cost, a[h0:h1].copy(), b[h0:h1].copy(), 203, 300, r0, r1, h0, h1, 5, dmin, dmax, torch.int16)
                torch.cuda.synchronize()
                assert np.array_equal(band.cpu().numpy(), full[r0:r1]), (cost, rank)
    sh = sharding.ShardedStereo(comp)
    out = sh.disparity_bands(sb.COST_SSD, L, Rt, 5, -39, 0, dtype=torch.int16)
    assert np.array_equal(out.cpu().numpy(), ctx.disparity(sb.COST_SSD, L, Rt, 5, -39, 0, dtype=np.int16))
    # both maps of every band of a 3-way split from one (fused) launch sequence each == the full-image pair call
    Lw, Rw, _ = synth.make_pair(101, 420, 64, 6)
    full_l, full_r = ctx.disparity_pair(sb.COST_SSD, Lw, Rw, 4, 127, dtype=np.int16)
    for rank in range(3):
        r0, r1 = sharding.band_shard(101, 3, rank)
        h0, h1 = sharding.band_halo(101, r0, r1, 4)
        both = comp.pair_band(sb.COST_SSD, Lw[h0:h1].copy(), Rw[h0:h1].copy(), 101, 420, r0, r1, h0, h1, 4, 127, torch.int16)
        torch.cuda.synchronize()
        assert ctx.last_fused_pairs == 1
        assert np.array_equal(both[0].cpu().numpy(), full_l[r0:r1]) and np.array_equal(both[1].cpu().numpy(), full_r[r0:r1]), rank
    pl, pr = sh.disparity_pair_bands(sb.COST_SSD, Lw, Rw, 4, 127, dtype=torch.int16)
    assert np.array_equal(pl.cpu().numpy(), full_l) and np.array_equal(pr.cpu().numpy(), full_r)
    Ls, Rs = np.stack([L[:64, :128], L[64:128, :128]]), np.stack([Rt[:64, :128], Rt[64:128, :128]])
    dl, dr = sh.disparity_pair_batch(sb.COST_SSD, Ls, Rs, 3, 15, dtype=torch.int8)
    bl, br = ctx.disparity_pair_batch(sb.COST_SSD, Ls, Rs, 3, 15)
    assert np.array_equal(dl.cpu().numpy(), bl) and np.array_equal(dr.cpu().numpy(), br)


def _peer_worker(rank, world, port, q):
    """Two ranks on the SAME device (the GPU test box has one B200): CUDA IPC works between processes on one
    device, so this exercises exactly the multi-GPU plumbing — handle exchange, peer mapping, copy-engine
    pushes ordered after the kernels, tickets, the completing barrier — with gloo as the process group."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
        ctx = sb.Context(0)
        comp = sharding.GpuCompute(ctx, dev)
        sh = sharding.ShardedStereo(comp, gather="peer")
        ok = True
        L, Rt, _ = synth.make_pair(45, 96, 12, 77)
        for (a, b, dmin, dmax) in ((L, Rt, -11, 0), (Rt, L, 0, 11)):
            full = sh.disparity_bands(sb.COST_SSD, a, b, 3, dmin, dmax, dtype=torch.int16).cpu().numpy()
            ok &= bool(np.array_equal(full, ctx.disparity(sb.COST_SSD, a, b, 3, dmin, dmax, dtype=np.int16)))
        Ls = np.stack([synth.make_pair(20, 64, 8, 300 + i)[0] for i in range(5)])
        Rs = np.stack([synth.make_pair(20, 64, 8, 300 + i)[1] for i in range(5)])
        dl, dr = sh.disparity_pair_batch(sb.COST_SSD, Ls, Rs, 2, 7, dtype=torch.int8)
        bl, br = ctx.disparity_pair_batch(sb.COST_SSD, Ls, Rs, 2, 7)
        ok &= bool(np.array_equal(dl.cpu().numpy(), bl) and np.array_equal(dr.cpu().numpy(), br))
        # raw PeerGather: pipelined pushes with tickets, every rank ends up with every rank's bytes
        n = 1 << 16
        pg = sharding.PeerGather(ctx, world * n)
        st = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st):
            for it in range(3):
                src = torch.full((n,), 10 * it + rank + 1, dtype=torch.uint8, device=dev)
                pg.push(rank * n, src.data_ptr(), n, st.cuda_stream)
                pg.wait(pg.mark(), st.cuda_stream)      # src may be recycled by the allocator after this point
            pg.complete(st.cuda_stream)
            got = pg.local_bytes(dev).view(world, n).cpu().numpy()
        ok &= all(bool((got[r] == 20 + r + 1).all()) for r in range(world))
        pg.close()
        sh.close()
        ctx.close()
        q.put((rank, ok))
    except Exception as exc:
        q.put((rank, f"{type(exc).__name__}: {exc}"))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_gpu_peer_gather_two_ranks_one_device():
    import torch.multiprocessing as mp
    mctx = mp.get_context("spawn")
    q = mctx.Queue()
    port = _free_port()
    procs = [mctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in results) == [0, 1]
    assert all(ok is True for _, ok in results), results
