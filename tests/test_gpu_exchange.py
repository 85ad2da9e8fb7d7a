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
