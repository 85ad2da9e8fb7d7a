"""Sequence parallelism for the quantized DiT step: one process per GPU, NCCL over NVLink/NVSwitch.

Replaces xfuser/yunchang's Ulysses path (ViDiT-Q/examples/Wan2.1/wan/distributed/xdit_context_parallel.py:66-192;
xfuser hybrid/attn_layer.py:160-206 = 4 separate c10d all-to-alls per block) with one packed q|k|v exchange and one
output exchange per attention, issued as grouped NCCL send/recv (`batch_isend_irecv`), no xfuser dependency.

Every stage of a block except attention is token-local (per-token activation scales make the quantized linears
exactly invariant to sequence sharding), so ranks own contiguous token chunks
`x[r*L/P:(r+1)*L/P]` (xdit_context_parallel.py:131-133) with weights replicated.

Head/query decomposition.  P = Pu * Pr with Pu = gcd-style largest divisor of P that divides the head count:
  rank r = h*Pu + g   owns head-group g (H/Pu heads) for the query range of replica h (the tokens of ranks
  h*Pu .. h*Pu+Pu-1) against ALL keys/values.
Pr == 1 is plain Ulysses (12-head Wan-1.3B at P in {1,2,4}; 40-head 14B at P in {2,4,8}).  Pr == 2 covers the
1.3B model on 8 GPUs, where 12 heads do not divide by 8 (the reference refuses that case,
quant_generate.py:451-452): Ulysses-4 x 2-way key/value replication — K,V of a head group go to both replicas.
NVSwitch gives every pair full bandwidth, so the extra K/V copy costs bytes, not hops.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def _largest_head_divisor(world, heads):
    best = 1
    for pu in range(1, world + 1):
        if world % pu == 0 and heads % pu == 0:
            best = pu
    return best


class SequenceParallel:
    def __init__(self, group=None, attention_core=None, pipeline_chunks=1, exchange="auto"):
        """exchange: how q|k|v and the attention output travel between the ranks.
          "peer"  one b200q_scatter_rows launch stores every destination's slice straight into that rank's receive buffer
                  over NVLink (one symmetric allocation per rank, peers mapped as plain device pointers), one signal-pad
                  barrier before the data is consumed: no staging copies, no NCCL call on the data path;
          "nccl"  grouped NCCL send/recv (`batch_isend_irecv`) with group-major staging copies;
          "auto"  "peer" on CUDA with the nccl backend when the symmetric allocation succeeds, else "nccl"
                  (`exchange_in_use` / `exchange_fallback` say which and why).
        pipeline_chunks > 1 (nccl exchange only): the heads of a rank's head group are exchanged and attended in that many
        chunks, so the NCCL transfer of chunk c+1 (and the return of chunk c-1) runs while the attention core works on
        chunk c (send/recv are asynchronous on NCCL's stream; only the consumer waits).  Results are bit-identical to the
        single-exchange path - attention is independent per head."""
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError(exchange)
        self.exchange, self.exchange_in_use, self.exchange_fallback = exchange, None, None
        self._peer = {}
        self._peer_calls = 0
        self.pipeline_chunks = int(pipeline_chunks)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.attention_core = attention_core          # (q,k,v,num_heads)->o ; default library SDPA
        self.bytes_sent = 0                            # per-rank payload counter (bench/DESIGN accounting)

    # ---- token sharding ------------------------------------------------------------------------------------
    def chunk(self, L):
        P = self.world_size
        if L % P != 0:
            raise ValueError(f"sequence length {L} must divide by the sequence-parallel size {P} "
                             "(the reference pads to a multiple of sp_size, text2video.py:170-172)")
        return L // P

    def shard_tokens(self, x):
        Lr = self.chunk(x.shape[0])
        off = self.rank * Lr
        return x[off:off + Lr], off, Lr

    def gather_tokens(self, y, L):
        """all_gather along tokens (xdit_context_parallel.py:142)."""
        y = y.contiguous()
        out = torch.empty((self.world_size * y.shape[0],) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
        dist.all_gather_into_tensor(out, y, group=self.group)
        return out[:L]

    # ---- attention exchange ------------------------------------------------------------------------------------
    def plan(self, num_heads):
        P = self.world_size
        Pu = _largest_head_divisor(P, num_heads)
        Pr = P // Pu
        return Pu, Pr, self.rank % Pu, self.rank // Pu

    def _exchange(self, sends, recvs):
        """grouped point-to-point = NCCL all-to-all with per-peer sizes; self-copy done locally."""
        ops = []
        for peer in range(self.world_size):
            s, r = sends[peer], recvs[peer]
            if peer == self.rank:
                if r is not None and s is not None:
                    r.copy_(s)
                continue
            if r is not None and r.numel() > 0:
                ops.append(dist.P2POp(dist.irecv, r, self._global(peer), group=self.group))
            if s is not None and s.numel() > 0:
                ops.append(dist.P2POp(dist.isend, s, self._global(peer), group=self.group))
                self.bytes_sent += s.numel() * s.element_size()
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _exchange_async(self, sends, recvs):
        """like _exchange, but returns the outstanding work handles instead of waiting on them"""
        ops = []
        for peer in range(self.world_size):
            s, r = sends[peer], recvs[peer]
            if peer == self.rank:
                if r is not None and s is not None:
                    r.copy_(s)
                continue
            if r is not None and r.numel() > 0:
                ops.append(dist.P2POp(dist.irecv, r, self._global(peer), group=self.group))
            if s is not None and s.numel() > 0:
                ops.append(dist.P2POp(dist.isend, s, self._global(peer), group=self.group))
                self.bytes_sent += s.numel() * s.element_size()
        return dist.batch_isend_irecv(ops) if ops else []

    def _global(self, peer):
        return peer if self.group is None else dist.get_global_rank(self.group, peer)

    def attention(self, q, k, v, num_heads):
        """q,k,v: this rank's tokens, all heads, [Lr, H*hd] -> attention output [Lr, H*hd] for the same tokens."""
        from . import model as M                       # late import: model imports nothing from here
        P = self.world_size
        core = self.attention_core or M.attention_bf16
        if P == 1:
            return core(q, k, v, num_heads)
        Lr, D = q.shape
        hd = D // num_heads
        Pu, Pr, g, h = self.plan(num_heads)
        Hg = num_heads // Pu
        W = Hg * hd
        if self.exchange == "peer" and not q.is_cuda:
            raise RuntimeError("SequenceParallel(exchange='peer'): the peer-memory exchange needs CUDA tensors")
        if self.exchange != "nccl" and q.is_cuda:
            peer = self._peer_buffers(Lr, W, P, Pu, q.dtype, q.device)
            if peer is not None:
                return self._attention_peer(q, k, v, core, peer, Lr, W, Pu, Pr, g, h, Hg)
        self.exchange_in_use = "nccl"
        C = self.pipeline_chunks
        if C > 1 and Hg % C == 0:
            return self._attention_pipelined(q, k, v, core, Lr, hd, Pu, Pr, g, h, Hg, C)
        # Group-major staging: x [Lr, Pu, W] -> [Pu, Lr, W], so the columns a destination needs are ONE contiguous chunk and
        # every receive lands in place: K / V of all ranks as [P, Lr, W] == [L, W], Q of my replica as [Pu, Lr, W].  Three
        # transposing copies in, one out, and two grouped NCCL send/recv batches per attention - instead of a stack per
        # destination and a cat per tensor.
        gm = lambda x: x.view(Lr, Pu, W).permute(1, 0, 2).contiguous()
        qg, kg, vg = gm(q), gm(k), gm(v)
        Kr = torch.empty((P, Lr, W), dtype=q.dtype, device=q.device)
        Vr = torch.empty((P, Lr, W), dtype=q.dtype, device=q.device)
        Qr = torch.empty((Pu, Lr, W), dtype=q.dtype, device=q.device)
        ops = []
        for peer in range(P):
            gp, hp = peer % Pu, peer // Pu
            if peer == self.rank:
                Kr[peer].copy_(kg[gp]); Vr[peer].copy_(vg[gp]); Qr[gp].copy_(qg[gp])
                continue
            ops.append(dist.P2POp(dist.irecv, Kr[peer], self._global(peer), group=self.group))
            ops.append(dist.P2POp(dist.irecv, Vr[peer], self._global(peer), group=self.group))
            if hp == h:
                ops.append(dist.P2POp(dist.irecv, Qr[gp], self._global(peer), group=self.group))
            ops.append(dist.P2POp(dist.isend, kg[gp], self._global(peer), group=self.group))
            ops.append(dist.P2POp(dist.isend, vg[gp], self._global(peer), group=self.group))
            self.bytes_sent += 2 * kg[gp].numel() * kg.element_size()
            if hp == h:                                # my tokens' queries are served by the ranks of my replica
                ops.append(dist.P2POp(dist.isend, qg[gp], self._global(peer), group=self.group))
                self.bytes_sent += qg[gp].numel() * qg.element_size()
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        O = core(Qr.view(Pu * Lr, W), Kr.view(P * Lr, W), Vr.view(P * Lr, W), Hg)   # [L/Pr, W]
        # return each token owner its rows of my head group
        Or = torch.empty((Pu, Lr, W), dtype=O.dtype, device=O.device)
        ops = []
        for i in range(Pu):
            peer = h * Pu + i                          # owner of query rows i*Lr .. (i+1)*Lr; it computed head group i for me
            if peer == self.rank:
                Or[i].copy_(O[i * Lr:(i + 1) * Lr])
                continue
            ops.append(dist.P2POp(dist.irecv, Or[i], self._global(peer), group=self.group))
            ops.append(dist.P2POp(dist.isend, O[i * Lr:(i + 1) * Lr], self._global(peer), group=self.group))
            self.bytes_sent += Lr * W * O.element_size()
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return Or.permute(1, 0, 2).reshape(Lr, Pu * W)                          # [Lr, H*hd]

    # ---- exchange over NVLink peer memory --------------------------------------------------------------------
    def _peer_buffers(self, Lr, W, P, Pu, dtype, device):
        """One symmetric allocation per rank and shape: [K of all ranks | V of all ranks | Q of my replica | 2 x output],
        plus every peer's base address.  None (with the reason in `exchange_fallback`) when it cannot be set up."""
        key = (Lr, W, P, Pu, dtype)
        if key in self._peer:
            return self._peer[key]
        try:
            if dist.get_backend(self.group) != "nccl":
                raise RuntimeError("peer-memory exchange needs CUDA ranks on one NVLink domain (backend %s)" % dist.get_backend(self.group))
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty(self.peer_numel(Lr, W, P, Pu), dtype=dtype, device=device)
            hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            peer = self.peer_layout(buf, Lr, W, P, Pu)
            peer["hdl"], peer["base"] = hdl, [int(p) for p in hdl.buffer_ptrs]
            hdl.barrier()
            self.exchange_in_use = "peer"
        except Exception as ex:                                   # no NVLink peer mapping here: stay on NCCL, say why
            if self.exchange == "peer":
                raise
            self.exchange_fallback = repr(ex)
            peer = None
        self._peer[key] = peer
        return peer

    def _peer_messages_in(self, q, k, v, peer, Lr, W, Pu, g, h):
        """(src, dst, src_pitch) of the operand exchange: my K / V columns of head group gp go to slot `my rank` of every
        rank computing gp; my Q columns of head group gp go to slot `my place in the replica` of the ranks of my replica."""
        P, es = self.world_size, q.element_size()
        base, off, me, slot = peer["base"], peer["off"], self.rank, Lr * W * q.element_size()
        src, dst, pitch = [], [], []
        for dest in range(P):
            gp, hp = dest % Pu, dest // Pu                    # the destination computes head group gp for replica hp's queries
            col = gp * W * es
            for name, t in (("k", k), ("v", v)):
                src.append(t.data_ptr() + col); dst.append(base[dest] + off[name] + me * slot); pitch.append(t.stride(0) * es)
            if hp == h:
                src.append(q.data_ptr() + col); dst.append(base[dest] + off["q"] + g * slot); pitch.append(q.stride(0) * es)
            if dest != me:
                self.bytes_sent += (3 if hp == h else 2) * slot
        return src, dst, pitch

    def _peer_messages_out(self, O, peer, okey, Lr, W, Pu, g, h):
        """(src, dst, src_pitch) of the output exchange: rows i*Lr.. of my result belong to rank h*Pu + i and land in column
        block g (the head group I computed) of its [Lr, Pu*W] output buffer."""
        es = O.element_size()
        base, off, me = peer["base"], peer["off"], self.rank
        src, dst, pitch = [], [], []
        for i in range(Pu):
            dest = h * Pu + i
            src.append(O.data_ptr() + i * Lr * O.stride(0) * es); dst.append(base[dest] + off[okey] + g * W * es)
            pitch.append(O.stride(0) * es)
            if dest != me:
                self.bytes_sent += Lr * W * es
        return src, dst, pitch

    def _attention_peer(self, q, k, v, core, peer, Lr, W, Pu, Pr, g, h, Hg):
        """attention() with both exchanges as direct stores into the peers' buffers (include/b200q.h b200q_scatter_rows).
        Layout at every destination: K / V of all ranks [P, Lr, W] == [L, W] (slot = sending rank), Q of the replica
        [Pu, Lr, W] (slot = the sender's place in the replica), output [Lr, Pu * W] (column block = the head group the
        sender computed) - exactly what the attention kernel and the output projection read, so nothing is repacked.
        Two barriers per attention: operands landed / outputs landed.  The second one also orders this block's reads of
        K, V, Q before the next block's stores into them (a rank only reaches it after its own attention kernel)."""
        import b200q
        es = q.element_size()
        src, dst, pitch = self._peer_messages_in(q, k, v, peer, Lr, W, Pu, g, h)
        b200q.scatter_rows(src, dst, Lr, W * es, pitch, W * es)
        peer["hdl"].barrier()
        O = core(peer["Q"], peer["K"], peer["V"], Hg)              # [Pu * Lr, W]: my head group for the replica's queries
        which = self._peer_calls & 1                                # two output buffers: CFG branches are attended back to back
        self._peer_calls += 1
        src, dst, pitch = self._peer_messages_out(O, peer, "o1" if which else "o0", Lr, W, Pu, g, h)
        b200q.scatter_rows(src, dst, Lr, W * es, pitch, Pu * W * es)
        peer["hdl"].barrier()
        return peer["O"][which]                                     # [Lr, H*hd]

    @staticmethod
    def peer_layout(buf, Lr, W, P, Pu):
        """Views and byte offsets of one rank's exchange buffer: [K of all ranks | V of all ranks | Q of my replica | 2 x output]."""
        n_k, n_q, n_o = P * Lr * W, Pu * Lr * W, Lr * Pu * W
        es = buf.element_size()
        off = {"k": 0, "v": n_k, "q": 2 * n_k, "o0": 2 * n_k + n_q, "o1": 2 * n_k + n_q + n_o}
        return {
            "buf": buf, "off": {k_: v_ * es for k_, v_ in off.items()},
            "K": buf[off["k"]:off["k"] + n_k].view(P * Lr, W), "V": buf[off["v"]:off["v"] + n_k].view(P * Lr, W),
            "Q": buf[off["q"]:off["q"] + n_q].view(Pu * Lr, W),
            "O": [buf[off["o0"]:off["o0"] + n_o].view(Lr, Pu * W), buf[off["o1"]:off["o1"] + n_o].view(Lr, Pu * W)],
        }

    @staticmethod
    def peer_numel(Lr, W, P, Pu):
        return 2 * P * Lr * W + Pu * Lr * W + 2 * Lr * Pu * W

    def _attention_pipelined(self, q, k, v, core, Lr, hd, Pu, Pr, g, h, Hg, C):
        """Head-chunked variant of attention(): the heads of a group travel and are attended in C chunks.  Staging layout
        [Pu, C, Lr, Wc] makes every (destination group, chunk) one contiguous message; all inbound exchanges are posted up
        front (NCCL runs them back to back on its own stream), each chunk's attention waits only for its own operands, and
        its output travels back while the next chunk is being computed."""
        P = self.world_size
        Hc = Hg // C
        Wc = Hc * hd
        gm = lambda x: x.view(Lr, Pu, C, Wc).permute(1, 2, 0, 3).contiguous()
        qg, kg, vg = gm(q), gm(k), gm(v)
        Kr = torch.empty((C, P, Lr, Wc), dtype=q.dtype, device=q.device)
        Vr = torch.empty((C, P, Lr, Wc), dtype=q.dtype, device=q.device)
        Qr = torch.empty((C, Pu, Lr, Wc), dtype=q.dtype, device=q.device)
        inbound = []
        for c in range(C):
            ops = []
            for peer in range(P):
                gp, hp = peer % Pu, peer // Pu
                if peer == self.rank:
                    Kr[c, peer].copy_(kg[gp, c]); Vr[c, peer].copy_(vg[gp, c]); Qr[c, gp].copy_(qg[gp, c])
                    continue
                ops.append(dist.P2POp(dist.irecv, Kr[c, peer], self._global(peer), group=self.group))
                ops.append(dist.P2POp(dist.irecv, Vr[c, peer], self._global(peer), group=self.group))
                if hp == h:
                    ops.append(dist.P2POp(dist.irecv, Qr[c, gp], self._global(peer), group=self.group))
                ops.append(dist.P2POp(dist.isend, kg[gp, c], self._global(peer), group=self.group))
                ops.append(dist.P2POp(dist.isend, vg[gp, c], self._global(peer), group=self.group))
                self.bytes_sent += 2 * Lr * Wc * q.element_size()
                if hp == h:
                    ops.append(dist.P2POp(dist.isend, qg[gp, c], self._global(peer), group=self.group))
                    self.bytes_sent += Lr * Wc * q.element_size()
            inbound.append(dist.batch_isend_irecv(ops) if ops else [])
        Or = torch.empty((Pu, Lr, C, Wc), dtype=q.dtype, device=q.device)        # [group, row, chunk, Wc] == [Pu, Lr, W]
        outbound, keep = [], []
        for c in range(C):
            for w in inbound[c]:
                w.wait()
            O = core(Qr[c].view(Pu * Lr, Wc), Kr[c].view(P * Lr, Wc), Vr[c].view(P * Lr, Wc), Hc)   # [L/Pr, Wc]
            Oc = torch.empty((Pu, Lr, Wc), dtype=O.dtype, device=O.device)       # what the peers computed for my rows
            ops = []
            for i in range(Pu):
                peer = h * Pu + i
                if peer == self.rank:
                    Oc[i].copy_(O[i * Lr:(i + 1) * Lr])
                    continue
                ops.append(dist.P2POp(dist.irecv, Oc[i], self._global(peer), group=self.group))
                ops.append(dist.P2POp(dist.isend, O[i * Lr:(i + 1) * Lr], self._global(peer), group=self.group))
                self.bytes_sent += Lr * Wc * O.element_size()
            outbound.append(dist.batch_isend_irecv(ops) if ops else [])
            keep.append((O, Oc))
        for c in range(C):
            for w in outbound[c]:
                w.wait()
            Or[:, :, c].copy_(keep[c][1])
        return Or.permute(1, 0, 2, 3).reshape(Lr, Pu * C * Wc)                   # [Lr, H*hd]: group, chunk, head-in-chunk

    # ---- calibration ---------------------------------------------------------------------------------------------
    def allreduce_max(self, flat_stats):
        """Merge per-rank running abs-max over the sequence-sharded tokens: ONE allreduce(MAX) on the flat fp32
        buffer instead of pickled all_gather_object + cat + max (get_calib_data_wanx.py:443-468, ptq_wanx.py:336)."""
        if self.world_size > 1:
            dist.all_reduce(flat_stats, op=dist.ReduceOp.MAX, group=self.group)
        return flat_stats


def exchange_bytes_per_rank(L, D, P, num_heads, elem=2):
    """Algorithmic payload one rank sends per attention (q|k|v out + o back), for the roofline notes."""
    Pu = _largest_head_divisor(P, num_heads)
    Pr = P // Pu
    Lr, W = L // P, D // Pu
    kv = 2 * Lr * W * elem * (P - 1)
    q = Lr * W * elem * (Pu - 1)
    o = Lr * W * elem * (Pu - 1)
    return kv + q + o, Pu, Pr
