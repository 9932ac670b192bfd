"""Ulysses-style head/sequence exchange around the ASA call (absent from the reference,
which only replicates whole pipelines per GPU, simple_multiprocess_sampler.py:296-339).

Outside attention a rank owns S/P contiguous tokens x all H heads; inside it owns all S tokens x H/P heads.
ASA is independent per (batch, head) -- estimator, selection, pooled branch and the Gilbert permutation all
act on one head's full sequence -- so head sharding is exact.  The exchange is one `all_to_all_single`
(NCCL over NVLink/NVSwitch; gloo in the CPU tests) per tensor each way.  Ranks are arranged as
`world // P` independent groups of P consecutive ranks (CFG / prompt batch split across groups: no traffic).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class UlyssesGroup:
    def __init__(self, world: int, rank: int, degree: int):
        assert world % degree == 0
        self.P = degree
        self.rank_in_group = rank % degree
        self.group_id = rank // degree
        self.group = None
        self._vrow = {}
        if degree > 1:
            for g in range(world // degree):            # every rank must create every group
                ranks = list(range(g * degree, (g + 1) * degree))
                pg = dist.new_group(ranks)
                if g == self.group_id:
                    self.group = pg

    # [B=1, S/P, H, D] (my tokens, all heads)  ->  [1, S, H/P, D] (all tokens, my heads)
    def scatter_heads(self, *xs):
        if self.P == 1:
            return xs if len(xs) > 1 else xs[0]
        outs = []
        for x in xs:
            B, Sl, H, D = x.shape
            assert B == 1 and H % self.P == 0
            Hl = H // self.P
            send = x.view(Sl, self.P, Hl, D).permute(1, 0, 2, 3).contiguous()      # [P, S/P, Hl, D]
            recv = torch.empty_like(send)
            dist.all_to_all_single(recv, send, group=self.group)                    # chunk p <- rank p's tokens
            outs.append(recv.view(1, self.P * Sl, Hl, D))
        return outs if len(outs) > 1 else outs[0]

    # Fused variant: q, k, v travel in ONE all_to_all and are never unpacked.  The receive buffer is
    # [P(src rank), 3, S/P, H/P, D]; token s = p*(S/P)+i of tensor j sits at "virtual row" p*3*(S/P) + j*(S/P) + i
    # (row = H/P*D elements).  Returned: strided q/k/v views whose row stride is one virtual row plus the int32
    # table virtual_row[s] for q (k, v views start S/P and 2*S/P rows later), to be composed with the Gilbert
    # gather table so that the prep kernel reads straight out of the receive buffer.
    def scatter_heads_fused(self, q, k, v):
        B, Sl, H, D = q.shape
        assert B == 1 and H % self.P == 0
        P, Hl = self.P, H // self.P
        if P == 1:
            # degree 1: nothing to exchange (and `self.group` is None, i.e. the WORLD group -- an all_to_all here
            # would mix the sequences of different CFG groups).  Same return contract: [1,H,S,D] views of the
            # caller's token-major tensors and the identity row table.
            key = (str(q.device), Sl)
            if key not in self._vrow:
                self._vrow[key] = torch.arange(Sl, dtype=torch.int32, device=q.device)
            return q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), self._vrow[key], (q, k, v)
        send = torch.empty(P, 3, Sl, Hl, D, dtype=q.dtype, device=q.device)
        for j, x in enumerate((q, k, v)):
            send[:, j].copy_(x.view(Sl, P, Hl, D).permute(1, 0, 2, 3))
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        S = P * Sl
        row = Hl * D
        views = [recv.as_strided((1, Hl, S, D), (0, D, row, 1), j * Sl * row) for j in range(3)]
        key = (str(q.device), Sl)
        if key not in self._vrow:
            s_idx = torch.arange(S, device=q.device)
            self._vrow[key] = ((s_idx // Sl) * (3 * Sl) + (s_idx % Sl)).to(torch.int32)
        return views[0], views[1], views[2], self._vrow[key], recv

    # [1, S, H/P, D] (all tokens, my heads)  ->  [1, S/P, H, D] (my tokens, all heads)
    def gather_heads(self, o):
        if self.P == 1:
            return o
        B, S, Hl, D = o.shape
        assert B == 1 and S % self.P == 0
        Sl = S // self.P
        send = o.contiguous().view(self.P, Sl, Hl, D)                               # chunk p -> rank p
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)                        # chunk p <- rank p's heads
        return recv.permute(1, 0, 2, 3).reshape(1, Sl, self.P * Hl, D)


class UlyssesPeerPlane:
    """The Ulysses exchange as loads and stores over NVLink peer memory instead of two all-to-alls (SURVEY.md 8e,
    include/blade_asa.h: BladePeers).  Per rank and group, symmetric buffers (torch.distributed._symmetric_memory:
    CUDA VMM allocations mapped into every peer of the group):

        qkv  [3, Sl, H, D]   my token shard's q/k/v projections (the caller writes them here, e.g. addmm(out=...))
        out  [Sl, H, D]      my token shard's attention output, every head -- written by the peers' epilogues
        rstd [2, S]          Wan q/k RMSNorm statistic of ALL tokens -- every rank stores its shard into all peers

    One layer = (rstd push) -> barrier -> blade_asa_forward(peers=...) -> barrier:
      * the gather kernel PULLS its heads' rows of q, k, v from the owning peers (no pack, no send buffer, no NCCL);
      * the attention epilogue PUSHES each finished 128-row tile to the peers that own those tokens while the tensor
        cores run the next tile -- the return all-to-all disappears into the kernel.
    The two barriers are `_SymmetricMemory.barrier` (a few-microsecond signal exchange on the current stream)."""

    def __init__(self, group: UlyssesGroup, Sl: int, H: int, D: int, dtype=torch.bfloat16, device=None):
        import torch.distributed._symmetric_memory as symm
        assert group.P > 1 and H % group.P == 0
        self.g, self.Sl, self.H, self.D = group, Sl, H, D
        self.P, self.Hl, self.S = group.P, H // group.P, Sl * group.P
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        esz = torch.empty(0, dtype=dtype).element_size()
        n_qkv, n_out, n_rstd = 3 * Sl * H * D * esz, Sl * H * D * esz, 2 * self.S * 4
        off_out = (n_qkv + 1023) // 1024 * 1024
        off_rstd = (off_out + n_out + 1023) // 1024 * 1024
        total = off_rstd + n_rstd
        self.buf = symm.empty(total, dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, group.group)
        assert self.hdl.world_size == self.P and self.hdl.rank == group.rank_in_group
        self.qkv = self.buf[:n_qkv].view(dtype).view(3, Sl, H, D)
        self.out = self.buf[off_out:off_out + n_out].view(dtype).view(Sl, H, D)
        self.rstd = self.buf[off_rstd:off_rstd + n_rstd].view(torch.float32).view(2, self.S)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        from ._lib import BladePeers
        import ctypes as C
        pe = BladePeers()
        pe.n_peers, pe.my_peer, pe.rows_per_peer = self.P, group.rank_in_group, Sl
        for p in range(self.P):
            pe.q[p] = ptrs[p]
            pe.k[p] = ptrs[p] + Sl * H * D * esz
            pe.v[p] = ptrs[p] + 2 * Sl * H * D * esz
            pe.out[p] = ptrs[p] + off_out
        self.peers = pe
        self._rstd_ptrs = (C.c_void_p * self.P)(*[ptrs[p] + off_rstd for p in range(self.P)])

    def barrier(self):
        self.hdl.barrier(channel=0)

    def views(self):
        """[1, Hl, S, D] layout descriptors of q/k/v for the engine: geometry only -- the S rows live in the P peers'
        buffers (Sl each) and are pulled from there, so no local tensor can back them."""
        from ._lib import TensorLayout
        Hl, S, D, H = self.Hl, self.S, self.D, self.H
        esz = self.qkv.element_size()
        base = self.g.rank_in_group * Hl * D * esz
        return tuple(TensorLayout(self.qkv[j].data_ptr() + base, (1, Hl, S, D), (S * H * D, D, H * D, 1), self.qkv.dtype,
                                  self.qkv.device) for j in range(3))

    def push_rms_stat(self, eng, eps: float):
        """rstd of my tokens (q and k, all heads: MW:99-102) stored into every peer's table."""
        import ctypes as C
        from ._lib import check, current_stream, tensor_desc
        q = self.qkv[0].unsqueeze(0).transpose(1, 2)                    # [1, H, Sl, D] view of token-major memory
        k = self.qkv[1].unsqueeze(0).transpose(1, 2)
        check(eng.lib.blade_qk_rms_stat_peers(C.byref(tensor_desc(q)), C.byref(tensor_desc(k)), float(eps),
                                              self._rstd_ptrs, self.P, self.S, self.g.rank_in_group * self.Sl,
                                              current_stream()))

    def launch_args(self):
        """(peers, out) for AsaEngine.forward.  BLADE_PEER_NO_PUSH=1 (diagnostic): the output stays local instead of
        being stored into the peers' buffers -- isolates what the NVLink stores cost the attention epilogue."""
        if not os.environ.get("BLADE_PEER_NO_PUSH"):
            return self.peers, None
        from ._lib import BladePeers
        pe = BladePeers()
        pe.n_peers, pe.my_peer, pe.rows_per_peer = self.peers.n_peers, self.peers.my_peer, self.peers.rows_per_peer
        for p in range(self.P):
            pe.q[p], pe.k[p], pe.v[p] = self.peers.q[p], self.peers.k[p], self.peers.v[p]
        if getattr(self, "_local_out", None) is None:
            self._local_out = torch.empty(1, self.S, self.Hl, self.D, dtype=self.qkv.dtype,
                                          device=self.qkv.device).transpose(1, 2)
        return pe, self._local_out

    def attention(self, eng, **kw):
        """barrier -> layer (pull q/k/v, push out) -> barrier.  Returns my [Sl, H, D] output buffer (valid on the
        current stream after the call)."""
        self.barrier()                                                   # every peer's q/k/v (and rstd) are written
        q, k, v = self.views()
        pe, out = self.launch_args()
        _, cnt = eng.forward(q, k, v, peers=pe, out=out, **kw)
        self.barrier()                                                   # every peer's output rows have landed
        return self.out, cnt
