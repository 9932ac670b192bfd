"""Ulysses-style head/sequence exchange around the ASA call (absent from the reference,
which only replicates whole pipelines per GPU, simple_multiprocess_sampler.py:296-339).

Outside attention a rank owns S/P contiguous tokens x all H heads; inside it owns all S tokens x H/P heads.
ASA is independent per (batch, head) -- estimator, selection, pooled branch and the Gilbert permutation all
act on one head's full sequence -- so head sharding is exact.  The exchange is one `all_to_all_single`
(NCCL over NVLink/NVSwitch; gloo in the CPU tests) per tensor each way.  Ranks are arranged as
`world // P` independent groups of P consecutive ranks (CFG / prompt batch split across groups: no traffic).
"""
from __future__ import annotations

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
