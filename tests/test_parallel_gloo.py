"""CPU suite, world_size 2 and 4 over gloo: the sequence-parallel attention exchange (wan_b200/parallel.py) reproduces
single-process attention exactly (pure permutation + the same attention core), including the P = Pu x Pr hybrid used
when the head count does not divide by P, and the calibration allreduce(MAX) equals the unsharded statistic."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_attention(q, k, v, heads):
    Lq, hd = q.shape[0], q.shape[1] // heads
    qh, kh, vh = (t.view(-1, heads, hd).permute(1, 0, 2).unsqueeze(0) for t in (q, k, v))
    o = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)
    return o.squeeze(0).permute(1, 0, 2).reshape(Lq, heads * hd)


def _worker(rank, world, port, heads, L, hd, out_q, chunks=1):
    sys.path.insert(0, os.path.join(ROOT, "wan2.1-quantization_b200"))
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wan_b200.parallel import SequenceParallel
    torch.manual_seed(0)
    D = heads * hd
    q, k, v = (torch.randn(L, D) for _ in range(3))
    full = _cpu_attention(q, k, v, heads)
    sp = SequenceParallel(attention_core=_cpu_attention, pipeline_chunks=chunks)
    (qs, off, Lr), (ks, _, _), (vs, _, _) = sp.shard_tokens(q), sp.shard_tokens(k), sp.shard_tokens(v)
    o = sp.attention(qs.contiguous(), ks.contiguous(), vs.contiguous(), heads)
    ok_attn = torch.allclose(o, full[off:off + Lr], rtol=1e-5, atol=1e-6)
    if chunks > 1:       # the head-chunked pipeline is the same permutation: bit-identical to the single exchange
        sp1 = SequenceParallel(attention_core=_cpu_attention)
        ok_attn = ok_attn and torch.equal(o, sp1.attention(qs.contiguous(), ks.contiguous(), vs.contiguous(), heads))
    gathered = sp.gather_tokens(o, L)
    ok_gather = torch.allclose(gathered, full, rtol=1e-5, atol=1e-6)
    # calibration: sharded rows + allreduce(MAX) == unsharded abs-max, bit for bit
    x = torch.randn(L, 48)
    stat = x[off:off + Lr].abs().max(dim=0)[0]
    sp.allreduce_max(stat)
    ok_cal = torch.equal(stat, x.abs().max(dim=0)[0])
    out_q.put((rank, ok_attn, ok_gather, ok_cal, sp.plan(heads)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,heads", [(2, 4), (4, 12), (4, 6), (2, 3)])
def test_sequence_parallel_attention_gloo(world, heads):
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, heads, 48, 8, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out_q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_attn, ok_gather, ok_cal, plan in res:
        assert ok_attn and ok_gather and ok_cal, (rank, ok_attn, ok_gather, ok_cal, plan)


@pytest.mark.parametrize("world,heads,chunks", [(2, 12, 3), (4, 12, 3), (4, 8, 2), (4, 6, 3)])
def test_pipelined_exchange_gloo(world, heads, chunks):
    """pipeline_chunks > 1 (exchange of head chunk c+1 overlapping the attention of chunk c) gives the same result."""
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, heads, 48, 8, out_q, chunks)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out_q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_attn, ok_gather, ok_cal, plan in res:
        assert ok_attn and ok_gather and ok_cal, (rank, ok_attn, ok_gather, ok_cal, plan)


def test_head_plan():
    from wan_b200.parallel import _largest_head_divisor, exchange_bytes_per_rank
    assert _largest_head_divisor(8, 12) == 4 and _largest_head_divisor(4, 12) == 4 and _largest_head_divisor(8, 40) == 8
    b, pu, pr = exchange_bytes_per_rank(75600, 5120, 8, 40)
    assert (pu, pr) == (8, 1)
    # 14B P=8 sends (P-1)/P * (L/P)*D*2 = 84.7 MB per tensor per rank
    per_tensor = (8 - 1) / 8 * (75600 / 8) * 5120 * 2
    assert abs(b - 4 * per_tensor) < 1e-6 * b


def test_peer_exchange_layout_and_message_plan_cpu():
    """Host side of the NVLink peer-memory exchange (SequenceParallel._peer_messages_in/_out, peer_layout): every rank's
    messages, applied with plain tensor copies on CPU buffers, assemble exactly the operands / outputs the NCCL-shaped path
    produces - for Ulysses (Pr = 1) and the Pu x Pr hybrid.  (The device side, b200q_scatter_rows, is tests/test_gpu_exchange.py.)"""
    sys.path.insert(0, os.path.join(ROOT, "wan2.1-quantization_b200"))
    from wan_b200.parallel import SequenceParallel
    for P, H, L, hd in ((2, 4, 16, 8), (4, 12, 32, 8), (8, 12, 32, 8), (3, 6, 12, 8)):
        D, Lr = H * hd, L // P
        torch.manual_seed(P)
        q, k, v = (torch.randn(L, D).to(torch.bfloat16) for _ in range(3))
        sps = []
        for r in range(P):
            sp = SequenceParallel.__new__(SequenceParallel)
            sp.rank, sp.world_size, sp.group, sp.bytes_sent, sp._peer_calls = r, P, None, 0, 0
            sps.append(sp)
        Pu, Pr, _, _ = sps[0].plan(H)
        W = (H // Pu) * hd
        bufs = [torch.zeros(SequenceParallel.peer_numel(Lr, W, P, Pu), dtype=torch.bfloat16) for _ in range(P)]
        peers = [dict(SequenceParallel.peer_layout(bufs[r], Lr, W, P, Pu), base=[b.data_ptr() for b in bufs]) for r in range(P)]
        flat = {b.data_ptr(): b for b in bufs}

        def copy(src_t, src_ptr, dst_ptr, pitch, dst_pitch, rows, row_bytes):
            """emulate one b200q_scatter_rows message with tensor ops (2-byte elements)"""
            base = max(p_ for p_ in flat if p_ <= dst_ptr)
            d0, n = (dst_ptr - base) // 2, row_bytes // 2
            s0 = (src_ptr - src_t.data_ptr()) // 2
            sflat = src_t.as_strided((src_t.untyped_storage().nbytes() // 2 - src_t.storage_offset(),), (1,))
            for r_ in range(rows):
                flat[base][d0 + r_ * dst_pitch // 2: d0 + r_ * dst_pitch // 2 + n] = sflat[s0 + r_ * pitch // 2: s0 + r_ * pitch // 2 + n]

        rows = lambda t, r: t[r * Lr:(r + 1) * Lr]
        for r, sp in enumerate(sps):
            _, _, g, h = sp.plan(H)
            qs, ks, vs = rows(q, r), rows(k, r), rows(v, r)
            src, dst, pitch = sp._peer_messages_in(qs, ks, vs, peers[r], Lr, W, Pu, g, h)
            owners = [ks, vs, qs]
            for s_, d_, p_ in zip(src, dst, pitch):
                t = next(t for t in owners if t.data_ptr() <= s_ < t.data_ptr() + t.numel() * 2 + 1 and (s_ - t.data_ptr()) // 2 < D)
                copy(t, s_, d_, p_, W * 2, Lr, W * 2)
        for r, sp in enumerate(sps):
            _, _, g, h = sp.plan(H)
            # what rank (h, g) must hold: K / V of ALL tokens and Q of replica h's tokens, columns of head group g
            cols = slice(g * W, (g + 1) * W)
            assert torch.equal(peers[r]["K"], k[:, cols]) and torch.equal(peers[r]["V"], v[:, cols])
            assert torch.equal(peers[r]["Q"], q[h * Pu * Lr:(h + 1) * Pu * Lr, cols])
            O = peers[r]["Q"].clone()                                   # stand-in for the attention result [Pu*Lr, W]
            src, dst, pitch = sp._peer_messages_out(O, peers[r], "o0", Lr, W, Pu, g, h)
            for s_, d_, p_ in zip(src, dst, pitch):
                copy(O, s_, d_, p_, Pu * W * 2, Lr, W * 2)
        for r in range(P):                                              # every token owner gets all head groups of its rows back
            assert torch.equal(peers[r]["O"][0], rows(q, r))
    sp = SequenceParallel.__new__(SequenceParallel)
    sp.exchange, sp.world_size, sp.rank, sp.group, sp.attention_core, sp.pipeline_chunks = "peer", 2, 0, None, _cpu_attention, 1
    with pytest.raises(RuntimeError):
        sp.attention(torch.zeros(4, 16), torch.zeros(4, 16), torch.zeros(4, 16), 2)   # peer exchange demanded on CPU tensors
