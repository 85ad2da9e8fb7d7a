"""-m gpu: b200q_scatter_rows (include/b200q.h) - the data movement of the sequence-parallel attention exchange
(replaces the all-to-alls of wan/distributed/xdit_context_parallel.py:149-192).  One GPU: the destinations are local
buffers; the N > 1 path (peer buffers + barriers) is checked on hardware by `bench.py --gpus N` against the unsharded
step (`verify.bit_equal`).  Byte moves: exact equality."""
import pytest
import torch

import b200q

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("Lr,Pu,W,P", [(4095, 4, 384, 8), (16380, 2, 768, 2), (7, 3, 128, 3), (1, 1, 128, 1)])
def test_scatter_rows_head_group_exchange_layout(dev, Lr, Pu, W, P):
    """q|k|v [Lr, Pu*W] (v as a column slice of a wider matrix) -> per-destination slots, as SequenceParallel lays them out"""
    g = torch.Generator(device="cuda").manual_seed(Lr + W)
    q = torch.randn(Lr, Pu * W, device=dev, generator=g).to(torch.bfloat16)
    qkv = torch.randn(Lr, 3 * Pu * W, device=dev, generator=g).to(torch.bfloat16)
    v = qkv[:, 2 * Pu * W:]
    es = 2
    Kdst = [torch.zeros(P, Lr, W, dtype=torch.bfloat16, device=dev) for _ in range(P)]     # one receive buffer per "rank"
    Vdst = [torch.zeros(P, Lr, W, dtype=torch.bfloat16, device=dev) for _ in range(P)]
    me = P - 1
    src, dst, pitch = [], [], []
    for dest in range(P):
        gp = dest % Pu
        src += [q.data_ptr() + gp * W * es, v.data_ptr() + gp * W * es]
        dst += [Kdst[dest][me].data_ptr(), Vdst[dest][me].data_ptr()]
        pitch += [q.stride(0) * es, v.stride(0) * es]
    b200q.scatter_rows(src, dst, Lr, W * es, pitch, W * es)
    for dest in range(P):
        gp = dest % Pu
        assert torch.equal(Kdst[dest][me], q[:, gp * W:(gp + 1) * W])
        assert torch.equal(Vdst[dest][me], v[:, gp * W:(gp + 1) * W])
        for other in range(P - 1):
            assert not Kdst[dest][other].any()


def test_scatter_rows_output_return_layout_and_errors(dev):
    """attention output rows i*Lr.. -> column block g of the owner's [Lr, Pu*W] buffer (destination pitch > row)"""
    Lr, Pu, W, gme = 300, 4, 256, 2
    g = torch.Generator(device="cuda").manual_seed(3)
    O = torch.randn(Pu * Lr, W, device=dev, generator=g).to(torch.bfloat16)
    outs = [torch.zeros(Lr, Pu * W, dtype=torch.bfloat16, device=dev) for _ in range(Pu)]
    src = [O.data_ptr() + i * Lr * W * 2 for i in range(Pu)]
    dst = [outs[i].data_ptr() + gme * W * 2 for i in range(Pu)]
    b200q.scatter_rows(src, dst, Lr, W * 2, [W * 2] * Pu, Pu * W * 2)
    for i in range(Pu):
        assert torch.equal(outs[i][:, gme * W:(gme + 1) * W], O[i * Lr:(i + 1) * Lr])
        assert not outs[i][:, :gme * W].any() and not outs[i][:, (gme + 1) * W:].any()
    b200q.scatter_rows([], [], Lr, W * 2, [], W * 2)                                     # no messages: a no-op
    with pytest.raises(b200q.B200QError):
        b200q.scatter_rows(src, dst, Lr, W * 2 + 2, [W * 2] * Pu, Pu * W * 2)             # row not a multiple of 16 bytes
    with pytest.raises(b200q.B200QError):
        b200q.scatter_rows([src[0] + 2], [dst[0]], Lr, W * 2, [W * 2], Pu * W * 2)        # misaligned source
    with pytest.raises(b200q.B200QError):
        b200q.scatter_rows(src * 13, dst * 13, Lr, W * 2, [W * 2] * 52, Pu * W * 2)       # more than 48 messages


@pytest.mark.parametrize("P,H,L", [(2, 12, 1024), (4, 12, 2048), (8, 12, 2048), (8, 40, 1024), (3, 12, 768)])
def test_peer_exchange_all_ranks_on_one_gpu_equals_unsharded_attention(dev, P, H, L):
    """The message plan of SequenceParallel._attention_peer (slots, column blocks, Pu x Pr head/query decomposition incl. the
    1.3B-on-8-GPUs case Pu=4 x Pr=2), run for ALL P ranks on one GPU: every rank's buffer is a local tensor, the barriers are
    the phase boundaries of the loop.  The assembled result must be bit-equal to attention on the unsharded q, k, v with the
    same kernel and no key splits (attention is independent per head and per query row)."""
    from wan_b200.parallel import SequenceParallel
    hd = 128
    D = H * hd
    g0 = torch.Generator(device="cuda").manual_seed(P * 100 + H)
    q, k = (torch.randn(L, D, device=dev, generator=g0).to(torch.bfloat16) for _ in range(2))
    qkv = torch.randn(L, 3 * D, device=dev, generator=g0).to(torch.bfloat16)
    v = qkv[:, 2 * D:]                                              # a column slice, as in the block
    Lr = L // P
    sps = []
    for r in range(P):
        sp = SequenceParallel.__new__(SequenceParallel)
        sp.rank, sp.world_size, sp.group, sp.bytes_sent, sp._peer_calls = r, P, None, 0, 0
        sps.append(sp)
    Pu, Pr, _, _ = sps[0].plan(H)
    Hg = H // Pu
    W = Hg * hd
    bufs = [torch.zeros(SequenceParallel.peer_numel(Lr, W, P, Pu), dtype=torch.bfloat16, device=dev) for _ in range(P)]
    peers = []
    for r in range(P):
        peer = SequenceParallel.peer_layout(bufs[r], Lr, W, P, Pu)
        peer["base"] = [b.data_ptr() for b in bufs]
        peers.append(peer)
    rows = lambda t, r: t[r * Lr:(r + 1) * Lr]
    for r, sp in enumerate(sps):                                    # phase 1: every rank stores its slices
        _, _, g, h = sp.plan(H)
        src, dst, pitch = sp._peer_messages_in(rows(q, r), rows(k, r), rows(v, r), peers[r], Lr, W, Pu, g, h)
        b200q.scatter_rows(src, dst, Lr, W * 2, pitch, W * 2)
    outs = []
    for r, sp in enumerate(sps):                                    # phase 2 (after barrier 1): attention, outputs back
        _, _, g, h = sp.plan(H)
        O = b200q.attn_bf16(peers[r]["Q"], peers[r]["K"], peers[r]["V"], Hg, n_splits=1)
        outs.append(O)
        src, dst, pitch = sp._peer_messages_out(O, peers[r], "o0", Lr, W, Pu, g, h)
        b200q.scatter_rows(src, dst, Lr, W * 2, pitch, Pu * W * 2)
    got = torch.cat([peers[r]["O"][0] for r in range(P)], 0)        # phase 3 (after barrier 2): token owners read their rows
    ref = b200q.attn_bf16(q, k, v.contiguous(), H, n_splits=1)
    assert torch.equal(got, ref)
    assert sps[0].bytes_sent > 0
