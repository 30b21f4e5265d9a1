"""Drop-in `GP` module: the reference's gp.py API on top of the sm_100a kernels.

Same class name, constructor, parameter (`lvs`), method names, argument order, shapes and detach
semantics as /root/reference/pysrc/faceplace/gp.py:11-133, so `from gppvae_b200 import GP` can stand
where `from gp import GP` stood in train_gppvae.py:11.  Underneath, the reference's op sequence
(U, B = U^T U + I, svd, inverse, U B^-1, two GEMMs per solve) is replaced by the Q-space algorithm of
SURVEY.md section 7.2:

    pass 1    GC = V^T [V | X]                       gpp_gram_vtz     (+ NCCL all-reduce when sharded)
    factor    B = I + (v0/vn) G = Lc Lc^T, Linv      gpp_factor       (cached, see `_FactorCache`)
    solve     W = (v0/vn) B^-1 C                     gpp_solve_w
    pass 2    Xb = (X - V W)/vn, nll, sum Xb^2       gpp_xb_nll
    extras    vbs, Vb                                gpp_vbs, gpp_vb

Row sharding: construct with `process_group=` (or call `shard_rows(group)`); X and V are then this
rank's rows, GC and sum Xb^2 are all-reduced over the group and every rank holds identical W.
"""
from __future__ import annotations

import weakref
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from ._lib import S_QUAD, S_XB2
from .vmod import KhatriRao


class LowRankFactor:
    """What `GP.U_UBi_Shb` returns in place of the reference's dense `U` and `UBi` (gp.py:24-38).

    The reference materialises two extra N x Q tensors only to hand them back to `GP.solve`
    (train_gppvae.py:235-236).  This handle carries V, vs and the factorisation instead; `dense()`
    materialises the reference tensor on request.
    """

    def __init__(self, kind: str, V: torch.Tensor, ldv: int, Q: int, Qtrue: int, vs: torch.Tensor,
                 fac: ops.Factorisation):
        self.kind, self.V, self.ldv, self.Q, self.Qtrue, self.vs, self.fac = kind, V, ldv, Q, Qtrue, vs, fac
        self.n = V.shape[0]
        self.shape = torch.Size((self.n, Qtrue))

    def dense(self) -> torch.Tensor:
        vs = self.vs.detach()
        r = torch.sqrt(vs[0] / vs[-1])
        if self.kind == "U":
            return r * self.V[:, : self.Qtrue]
        if self.fac.Binv is None:
            raise RuntimeError("UBi.dense() needs B^-1: call GP.U_UBi_Shb(Vs, vs, want_binv=True)")
        zeros = torch.zeros(self.n, self.Q, device=self.V.device, dtype=torch.float32)
        VB = ops.x_minus_am(zeros, self.Q, self.V, self.ldv, self.fac.Binv, self.Q, self.n, self.Q, self.Q, -1.0)
        return r * VB[:, : self.Qtrue]


class KhatriRaoFactor:
    """`U` / `UBi` of `GP.U_UBi_Shb` when V was given in factored form (vmod.KhatriRao): carries the factors and the
    factorisation; `GP.solve` takes the structured route with it.  `dense()` goes through the dense handle."""

    def __init__(self, kind: str, kr: KhatriRao, vs: torch.Tensor, fac: ops.Factorisation):
        self.kind, self.kr, self.vs, self.fac = kind, kr, vs, fac
        self.n = kr.n
        self.shape = kr.shape

    def dense(self) -> torch.Tensor:
        Vm, ldv = ops.as_matrix(self.kr.dense(), "V")
        if Vm.shape[1] != self.fac.Q:       # p was padded to a multiple of 4: the extra columns of V are zero
            Vp = torch.zeros(self.n, self.fac.Q, device=Vm.device, dtype=torch.float32)
            Vp[:, : Vm.shape[1]] = Vm
            Vm, ldv = Vp, self.fac.Q
        return LowRankFactor(self.kind, Vm, ldv, self.fac.Q, self.shape[1], self.vs, self.fac).dense()


class LazyVb:
    """`Vbs[0] = dNLL/dV` (gp.py:68-71) of a structured evaluation: an N x Q matrix nobody needs whole -- the trainer only
    gathers minibatch rows from it (train_gppvae.py:283).  `[idx]` computes those rows, r L V[idx] B^-1 - Xb[idx] W^T."""

    def __init__(self, kr: KhatriRao, Xb, fac, W, scal, Q, Lk, L):
        self.kr, self.Xb, self.fac, self.W, self.scal, self.Q, self.Lk, self.L = kr, Xb, fac, W, scal, Q, Lk, L
        self.shape = kr.shape

    def __getitem__(self, idx) -> torch.Tensor:
        if not torch.is_tensor(idx):
            idx = torch.as_tensor(idx, device=self.kr.device)
        idx = idx.to(self.kr.device).reshape(-1)
        Vr = self.kr[idx]
        m = Vr.shape[0]
        Vp = torch.zeros(m, self.Q, device=Vr.device, dtype=torch.float32)
        Vp[:, : Vr.shape[1]] = Vr
        Xbr = self.Xb[idx].contiguous()
        Vb = ops.vb(Vp, self.Q, Xbr, self.fac.Binv, self.W, self.scal, m, self.Q, self.Lk, self.L)
        return Vb[:, : self.shape[1]]

    def dense(self, chunk: int = 65536) -> torch.Tensor:
        n = self.kr.n
        out = torch.empty(n, self.shape[1], device=self.kr.device, dtype=torch.float32)
        for a in range(0, n, chunk):
            idx = torch.arange(a, min(n, a + chunk), device=self.kr.device)
            out[a:a + idx.numel()] = self[idx]
        return out


class LazySingularValues:
    """`Shb` of gp.py:33 (singular values of B, descending).  The trainer discards it
    (train_gppvae.py:235), so it is only computed if somebody asks -- off the hot path, from G."""

    def __init__(self, G: torch.Tensor, Qtrue: int, vs: torch.Tensor):
        self._G, self._Q, self._vs, self._val = G, Qtrue, vs, None

    def value(self) -> torch.Tensor:
        if self._val is None:
            vs = self._vs.detach()
            B = torch.eye(self._Q, device=self._G.device) + (vs[0] / vs[-1]) * self._G[: self._Q, : self._Q]
            self._val = torch.linalg.eigvalsh(B.double()).flip(0).to(torch.float32)
        return self._val

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        args = tuple(a.value() if isinstance(a, LazySingularValues) else a for a in args)
        return func(*args, **(kwargs or {}))


class _FactorCache:
    """One-entry cache of (G, factorisation) keyed on the identity and version of V and of the variances.

    train_gppvae.py factors the same (Vt, vs) twice per epoch (:235 via U_UBi_Shb, then :166 inside
    taylor_coeff).  What keeps V's address from being recycled while the entry is alive: for matrices that have operand
    planes, the planes registry (`ops.PLANES`) -- the entry remembers WHICH planes object it was built from and is valid
    only while the registry still returns that very object for V (so dropping the registry entry, which
    `Vmodel.forward` does before it allocates the next V of the same shape, also retires this entry, and no second
    reference keeps last epoch's 16 GB alive); for small matrices the entry holds V itself.  In-place updates of V or lvs
    bump their version counters and miss; the variances are compared value for value as well.
    """

    def __init__(self):
        self.key = None
        self.keep = None
        self.token = None
        self.G = None
        self.ldg = 0
        self.C = None                # V^T X of the evaluation that built the entry (bench.run_check reads it)
        self.fac: Optional[ops.Factorisation] = None

    def lookup(self, key, want_binv: bool, vs: torch.Tensor, token=None):
        if self.key is not None and self.key == key and token is self.token and \
                (self.fac.Binv is not None or not want_binv):
            # version counters do not see writes through `.data` (the reference itself initialises parameters that way,
            # vmod.py:37-40): a hit also needs the variances the factorisation was built with, value for value
            if torch.equal(self.vs_snapshot, vs.detach().to(torch.float32)):
                return self.fac
        return None

    def store(self, key, keep, G, ldg, fac, vs: torch.Tensor, token=None):
        self.key, self.keep, self.G, self.ldg, self.fac, self.token = key, keep, G, ldg, fac, token
        self.vs_snapshot = vs.detach().to(torch.float32).clone()

    def clear(self):
        self.key = self.keep = self.G = self.C = self.fac = self.token = None


class GP(nn.Module):
    def __init__(self, n_rand_effs: int = 1, vsum2one: bool = True, process_group=None):
        super().__init__()
        if not vsum2one:
            # the reference's other branch is broken (gp.py:52 uses an undefined name)
            raise NotImplementedError("vsum2one=False is not implemented (it raises NameError in the reference too)")
        if n_rand_effs != 1:
            # the reference trainer only ever builds GP(n_rand_effs=1) (train_gppvae.py:138)
            raise NotImplementedError("only one random-effect design (n_rand_effs=1) is implemented")
        self.n_rand_effs = n_rand_effs
        self.vsum2one = vsum2one
        self.lvs = nn.Parameter(torch.zeros([n_rand_effs + 1]))      # gp.py:21-22
        self._group = process_group
        self._sharded = process_group is not None
        self._cache = _FactorCache()
        self._n_total_given = None
        self._vs_ref = None          # weakref to the tensor last returned by get_vs()
        self._vs_version = -1
        self.cache_hits = 0
        self.stage_hook = None       # optional callable(name): bench.py records CUDA events at stage boundaries
        self.overlap_exchange = True  # row shards: all-reduce G beside V^T X, C beside the Cholesky (see _factorise)

    # ------------------------------------------------------------------ sharding
    def shard_rows(self, process_group=None, n_total: Optional[int] = None) -> "GP":
        """Treat the rows handed to every method as this rank's shard of the N rows.

        `n_total` is the global number of rows.  Pass it when it is known (it usually is: the size of the data set):
        otherwise every evaluation all-reduces its local row count first, which costs a host synchronisation."""
        import torch.distributed as dist
        self._group = process_group if process_group is not None else dist.group.WORLD
        self._sharded = True
        self._n_total_given = n_total
        return self

    def invalidate_cache(self) -> None:
        """Forget the cached factorisation and operand planes.  Needed only after writes the version counters cannot
        see (`V.data[...] = ...`); in-place tensor operations and optimiser steps are detected."""
        self._cache.clear()
        ops.PLANES.invalidate()

    def _all_reduce(self, t: torch.Tensor, async_op: bool = False):
        """SUM over the row shards.  `async_op=True` returns the collective's work handle (None when not sharded): the
        caller keeps launching on its stream and calls `.wait()` where the result is first needed."""
        if self._sharded:
            import torch.distributed as dist
            return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._group, async_op=async_op)
        return None

    def _n_total(self, n: int, device) -> int:
        if not self._sharded:
            return n
        if self._n_total_given is not None:
            return int(self._n_total_given)
        # every rank takes part in this collective on every call (a per-rank cache keyed on the local row count would
        # let ranks disagree about whether to communicate)
        t = torch.tensor([n], device=device, dtype=torch.int64)
        self._all_reduce(t)
        return int(t.item())

    def _stage(self, name: str) -> None:
        if self.stage_hook is not None:
            self.stage_hook(name)

    # ------------------------------------------------------------------ reference API
    def get_vs(self) -> torch.Tensor:
        """softmax(lvs) = [v0, vn]  (gp.py:48-50)."""
        vs = F.softmax(self.lvs, 0)
        self._vs_ref, self._vs_version = weakref.ref(vs), self.lvs._version
        return vs

    def _vs_token(self, vs: torch.Tensor):
        """Cache token for a variance vector: tensors that came out of get_vs() for the current lvs are
        all the same value; anything else is identified by object and version (and kept alive)."""
        mine = self._vs_ref() if self._vs_ref is not None else None
        if mine is not None and vs is mine and self._vs_version == self.lvs._version:
            return ("lvs", id(self.lvs), self.lvs._version), None
        return ("tensor", id(vs), vs._version), vs

    def _cat(self, Vs: Sequence[torch.Tensor], vs: torch.Tensor) -> Tuple[torch.Tensor, int, int, int]:
        """gp.py:27 for the single design the trainer uses: V passes through untouched (the sqrt(vs[0])
        weight folds into the v0/vn factor inside the kernels).  Returns (V, ldv, Q padded, Q)."""
        if len(Vs) != 1:
            raise NotImplementedError(f"expected one design matrix in Vs, got {len(Vs)}")
        Vm, ldv = ops.as_matrix(Vs[0], "V")
        return Vm, ldv, Vm.shape[1], Vs[0].shape[1]

    def _factorise(self, Vm, ldv, Q, vs, want_binv: bool, Xm=None, ldx=0, Lk=0, vs_origin=None):
        """Pass 1 (+ all-reduce) and the factorisation, through the cache.
        Returns (fac, C, pV) with C = V^T X (Q x Lk, summed over ranks) or None when no X was given, and pV the operand
        planes of V (None for shapes below the tensor-core tile, which run on the fp32 engine)."""
        n = Vm.shape[0]
        tok, keep_vs = self._vs_token(vs if vs_origin is None else vs_origin)
        key = (Vm.untyped_storage().data_ptr(), Vm.storage_offset(), Vm._version, n, Q, ldv, tok)
        use_planes = ops.planes_supported(n, Q, Lk)
        pV = ops.planes_of(Vm, ldv) if use_planes else None

        def vtx():
            if use_planes:
                return ops.atb_planes(pV, ops.split_planes(Xm, ldx, n, Lk), n, Q, Lk)
            return ops.atb(Vm, ldv, Xm, ldx, n, Q, Lk)

        fac = self._cache.lookup(key, want_binv, vs, token=pV)
        if fac is not None:
            self.cache_hits += 1
            C = None
            if Lk:
                C = vtx()
                self._all_reduce(C)
            return fac, C, pV
        if self._sharded and Lk and self.overlap_exchange:
            # Row shards: the exchange step rides beside the work that does not need it.  Gram tiles -> all-reduce of G on
            # the collective's own stream WHILE this stream splits X and forms V^T X -> all-reduce of the small C WHILE
            # the Cholesky runs.  Only the tail of the first collective and nothing of the second stay exposed.
            self._stage("pass1:start")
            G = ops.gram_vtz_planes(pV, None, n, Q, 0) if use_planes else ops.gram_vtz(Vm, ldv, None, 0, n, Q, 0)
            hG = self._all_reduce(G, async_op=True)
            C = vtx()
            self._stage("pass1:end")
            hC = self._all_reduce(C, async_op=True)
            hG.wait()
            self._stage("allreduce:end")
            fac = ops.factor(G, Q, Q, vs, want_binv)
            hC.wait()
            self._stage("factor:end")
            self._cache.store(key, (None if use_planes else Vm, keep_vs), G, Q, fac, vs, token=pV)
            self._cache.C = C
            return fac, C, pV
        pX = ops.split_planes(Xm, ldx, n, Lk) if (use_planes and Lk) else None
        self._stage("pass1:start")
        GC = ops.gram_vtz_planes(pV, pX, n, Q, Lk) if use_planes else ops.gram_vtz(Vm, ldv, Xm, ldx, n, Q, Lk)
        self._stage("pass1:end")
        self._all_reduce(GC)
        self._stage("allreduce:end")
        fac = ops.factor(GC, Q + Lk, Q, vs, want_binv)
        self._stage("factor:end")
        # with planes the registry entry pins V (see _FactorCache); without, the cache entry does
        self._cache.store(key, (None if use_planes else Vm, keep_vs), GC, Q + Lk, fac, vs, token=pV)
        self._cache.C = GC[:, Q:] if Lk else None
        return fac, (GC[:, Q:] if Lk else None), pV

    # ------------------------------------------------------------------ structured route (vmod.KhatriRao)
    def _kr_c(self, kr: KhatriRao, Xm, ldx, Lk) -> torch.Tensor:
        """C = V^T X (Q x Lk) through the slot sums, summed over the ranks."""
        ST = kr.st(Xm, ldx, Lk, False)
        self._all_reduce(ST)
        return ops.kr_assemble_gc(ST, kr.wn, kr.p, Lk, False)

    def _kr_factorise(self, kr: KhatriRao, vs, want_binv: bool, Xm=None, ldx=0, Lk=0, vs_origin=None):
        """Structured pass 1 (+ all-reduce of the small ST instead of GC) and the factorisation, through the cache."""
        tok, keep_vs = self._vs_token(vs if vs_origin is None else vs_origin)
        key = ("kr", id(kr), tok)
        Q = kr.p * kr.q
        fac = self._cache.lookup(key, want_binv, vs)
        if fac is not None:
            self.cache_hits += 1
            return fac, (self._kr_c(kr, Xm, ldx, Lk) if Lk else None)
        if not Lk:      # factorisation alone (U_UBi_Shb): the slot sums still need a right-hand side to walk
            Xm = torch.zeros(kr.n, 4, device=kr.device, dtype=torch.float32)
            ldx, Lx = 4, 4
        else:
            Lx = Lk
        kr.index()
        self._stage("pass1:start")
        ST = kr.st(Xm, ldx, Lx, True)
        self._stage("pass1:end")
        self._all_reduce(ST)
        self._stage("allreduce:end")
        GC = ops.kr_assemble_gc(ST, kr.wn, kr.p, Lx, True)
        fac = ops.factor(GC, Q + Lx, Q, vs, want_binv)
        self._stage("factor:end")
        self._cache.store(key, (kr, keep_vs), GC, Q + Lx, fac, vs)
        return fac, (GC[:, Q:] if Lk else None)

    def _kr_solve(self, kr: KhatriRao, fac, C, Xm, ldx, Lk, L, n_total):
        """W, the scalar block, Xb = (X - V W)/vn and nll for the rows of X, V in factored form."""
        W, scal = ops.solve_w(fac, C, C.stride(0), Lk, L, n_total)
        self._stage("solve:end")
        M = ops.kr_assemble_m(W, kr.wn, kr.p, Lk)
        if ops.planes_supported(kr.P, kr.p, 0):
            pM = ops.split_planes(M, M.stride(0), kr.p, M.shape[1])
            Y = ops.am_planes(kr.xn_planes(), pM, kr.P, kr.p, M.shape[1])
        else:
            Y = ops.am(kr.xn, kr.p, M, M.stride(0), kr.P, kr.p, M.shape[1])
        Xb, nll = ops.kr_xb_nll(Xm, ldx, Y, kr.d, kr.w, kr.P, kr.nviews, Lk, scal)
        self._stage("pass2:end")
        return W, scal, Xb, nll

    def U_UBi_Shb(self, Vs: Sequence[torch.Tensor], vs: torch.Tensor, want_binv: bool = False):
        """gp.py:24-38.  Returns handles (see LowRankFactor) that `solve` accepts."""
        if len(Vs) == 1 and isinstance(Vs[0], KhatriRao):
            kr = Vs[0]
            fac, _ = self._kr_factorise(kr, vs, want_binv)
            return (KhatriRaoFactor("U", kr, vs, fac), KhatriRaoFactor("UBi", kr, vs, fac),
                    LazySingularValues(self._cache.G, kr.p * kr.q, vs))
        Vm, ldv, Q, Qtrue = self._cat(Vs, vs)
        fac, _, _ = self._factorise(Vm, ldv, Q, vs, want_binv)
        U = LowRankFactor("U", Vm, ldv, Q, Qtrue, vs, fac)
        UBi = LowRankFactor("UBi", Vm, ldv, Q, Qtrue, vs, fac)
        return U, UBi, LazySingularValues(self._cache.G, Qtrue, vs)

    def solve(self, X: torch.Tensor, U, UBi, vs: torch.Tensor) -> torch.Tensor:
        """K^-1 X  (gp.py:40-46)."""
        Xm, ldx = ops.as_matrix(X, "X")
        n, L = X.shape
        Lk = Xm.shape[1]
        if isinstance(U, KhatriRaoFactor):
            if n != U.n:
                raise ValueError(f"X has {n} rows but the factorisation was built for {U.n}")
            C = self._kr_c(U.kr, Xm, ldx, Lk)
            _, _, Xb, _ = self._kr_solve(U.kr, U.fac, C, Xm, ldx, Lk, L, self._n_total(n, X.device))
        elif isinstance(U, LowRankFactor):
            if n != U.n:
                raise ValueError(f"X has {n} rows but the factorisation was built for {U.n}")
            if ops.planes_supported(n, U.Q, Lk):
                pV = ops.planes_of(U.V, U.ldv)
                C = ops.atb_planes(pV, ops.split_planes(Xm, ldx, n, Lk), n, U.Q, Lk)
            else:
                pV = None
                C = ops.atb(U.V, U.ldv, Xm, ldx, n, U.Q, Lk)
            self._all_reduce(C)
            W, scal = ops.solve_w(U.fac, C, Lk, Lk, L, self._n_total(n, X.device))
            if pV is not None:
                Xb, _ = ops.xb_nll_planes(pV, Xm, ldx, W, n, U.Q, Lk, scal)
            else:
                Xb, _ = ops.xb_nll(U.V, U.ldv, Xm, ldx, W, n, U.Q, Lk, scal)
        else:
            # dense tensors, exactly gp.py:42-44: (X - UBi (U^T X)) / vn
            Um, ldu = ops.as_matrix(U, "U")
            UBm, ldub = ops.as_matrix(UBi, "UBi")
            UX = ops.atb(Um, ldu, Xm, ldx, n, Um.shape[1], Lk)
            self._all_reduce(UX)
            Xb = ops.x_minus_am(Xm, ldx, UBm, ldub, UX, Lk, n, Um.shape[1], Lk, 1.0 / float(vs.detach()[-1]))
        return Xb[:, :L] if Lk != L else Xb

    def _coefficients(self, X: torch.Tensor, Vs: Sequence[torch.Tensor], want_vb: bool):
        """Shared body of taylor_coeff and nll: returns a dict of everything computed."""
        vs_attached = self.get_vs()
        vs = vs_attached.detach()
        if len(Vs) == 1 and isinstance(Vs[0], KhatriRao):
            kr = Vs[0]
            Xm, ldx = ops.as_matrix(X, "X")
            n, L = X.shape
            if kr.n != n:
                raise ValueError(f"X has {n} rows but V has {kr.n}")
            if kr.device != Xm.device:
                raise ValueError("X and V must be on the same device")
            Lk = Xm.shape[1]
            n_total = self._n_total(n, X.device)
            fac, C = self._kr_factorise(kr, vs, want_vb, Xm, ldx, Lk, vs_origin=vs_attached)
            W, scal, Xb, nll = self._kr_solve(kr, fac, C, Xm, ldx, Lk, L, n_total)
            return dict(vs=vs, Vm=None, kr=kr, ldv=0, Q=kr.p * kr.q, Qtrue=kr.p_true * kr.q, n=n, L=L, Lk=Lk,
                        n_total=n_total, fac=fac, W=W, scal=scal, Xb=Xb, nll=nll)
        Vm, ldv, Q, Qtrue = self._cat(Vs, vs)
        Xm, ldx = ops.as_matrix(X, "X")
        n, L = X.shape
        if Vm.shape[0] != n:
            raise ValueError(f"X has {n} rows but V has {Vm.shape[0]}")
        if Vm.device != Xm.device:
            raise ValueError("X and V must be on the same device")
        Lk = Xm.shape[1]
        n_total = self._n_total(n, X.device)
        fac, C, pV = self._factorise(Vm, ldv, Q, vs, want_vb, Xm, ldx, Lk, vs_origin=vs_attached)
        W, scal = ops.solve_w(fac, C, C.stride(0), Lk, L, n_total)
        self._stage("solve:end")
        if pV is not None:
            Xb, nll = ops.xb_nll_planes(pV, Xm, ldx, W, n, Q, Lk, scal)
        else:
            Xb, nll = ops.xb_nll(Vm, ldv, Xm, ldx, W, n, Q, Lk, scal)
        self._stage("pass2:end")
        return dict(vs=vs, Vm=Vm, ldv=ldv, Q=Q, Qtrue=Qtrue, n=n, L=L, Lk=Lk, n_total=n_total, fac=fac, W=W,
                    scal=scal, Xb=Xb, nll=nll, pV=pV)

    def taylor_coeff(self, X: torch.Tensor, Vs: Sequence[torch.Tensor], need_vb: bool = True
                     ) -> Tuple[torch.Tensor, List[torch.Tensor], torch.Tensor, torch.Tensor]:
        """gp.py:55-95: Xb = dNLL/dX, [Vb = dNLL/dV], vbs = dNLL/d[v0, vn], nll (n x 1); all detached.

        `need_vb=False` (an extension) skips B^-1 and the N x Q matrix Vb and returns `[None]` in its place:
        that is the "NLL + dNLL/dZ" evaluation BASELINE.json's metric counts."""
        c = self._coefficients(X, Vs, need_vb)
        scal, Q, L, Lk = c["scal"], c["Q"], c["L"], c["Lk"]
        if self._sharded:
            self._all_reduce(scal[S_XB2:S_QUAD + 1])
        vbs = ops.vbs_from_scal(scal, c["n_total"], Q, L)
        if need_vb and c["Vm"] is None:
            Vbs = [LazyVb(c["kr"], c["Xb"], c["fac"], c["W"], scal, Q, Lk, L)]
        elif need_vb:
            if c.get("pV") is not None:
                Vb = ops.vb_planes(c["pV"], c["Xb"], c["fac"].Binv, c["W"], scal, c["n"], Q, Lk, L)
            else:
                Vb = ops.vb(c["Vm"], c["ldv"], c["Xb"], c["fac"].Binv, c["W"], scal, c["n"], Q, Lk, L)
            Vbs = [Vb[:, : c["Qtrue"]] if c["Qtrue"] != Q else Vb]
            self._stage("vb:end")
        else:
            Vbs = [None]
        self.last_scalars = scal
        Xb = c["Xb"][:, :L] if Lk != L else c["Xb"]
        return Xb, Vbs, vbs, c["nll"]

    def nll(self, X: torch.Tensor, Vs: Sequence[torch.Tensor]) -> torch.Tensor:
        """gp.py:97-110.  Differentiable: backward applies the Taylor coefficients (the exact gradients
        of sum(nll), gp.py:185-221), i.e. it is exact for a uniform upstream gradient (`.sum()`,
        `.mean()`) -- the only way the reference ever differentiates it; any other upstream gradient raises."""
        return _NllFunction.apply(self, X, self.lvs, *Vs)

    def nll_ineff(self, X: torch.Tensor, Vs: Sequence[torch.Tensor]) -> torch.Tensor:
        """gp.py:112-125: O(N^3) cross-check through the dense N x N covariance.  Test utility, stock torch."""
        vs = self.get_vs()
        V = torch.cat([torch.sqrt(vs[i]) * Vi for i, Vi in enumerate(Vs)], 1)
        K = V.mm(V.t()) + vs[-1] * torch.eye(X.shape[0], device=X.device, dtype=X.dtype)
        quad = (X * torch.linalg.solve(K, X)).sum(1, keepdim=True)
        return 0.5 * quad + 0.5 * X.shape[1] * torch.linalg.slogdet(K)[1] / X.shape[0]

    def taylor_expansion(self, X: torch.Tensor, Vs: Sequence[torch.Tensor], Xb: torch.Tensor,
                         Vbs: Sequence[torch.Tensor], vbs: torch.Tensor) -> torch.Tensor:
        """gp.py:127-133: (n x 1) surrogate with gradients to X, each V and lvs."""
        if len(Vs) != 1 or len(Vbs) != 1:
            raise NotImplementedError(f"expected one design matrix in Vs / Vbs, got {len(Vs)} / {len(Vbs)}")
        return ops._TaylorExpansion.apply(X, Vs[0], self.lvs, Xb, Vbs[0], vbs)


class _NllFunction(torch.autograd.Function):
    """gp.py:97-110 with the Taylor coefficients as its backward.

    The coefficients are the exact gradients of sum_i nll_i (gp.py:185-221).  nll_i couples all rows through K^-1, so a
    NON-uniform upstream gradient (a weighted sum of the rows) has a different gradient, which this class does not
    implement: it raises instead of returning the uniform-weight one.  `.sum()`, `.mean()` and any constant multiple --
    every way the reference ever differentiates nll -- are exact."""

    @staticmethod
    def forward(ctx, gp: GP, X, lvs, *Vs):
        want_grad = any(ctx.needs_input_grad[1:])
        dense = [torch.is_tensor(V) for V in Vs]
        ctx.dense = dense
        with torch.no_grad():
            if want_grad:
                # a factored V (vmod.KhatriRao) is a detached snapshot: no gradient flows to it, its Vb is never formed
                need_vb = any(d and ng for d, ng in zip(dense, ctx.needs_input_grad[3:]))
                Xb, Vbs, vbs, nll = gp.taylor_coeff(X, list(Vs), need_vb=need_vb)
                saved = [Vb for Vb, d in zip(Vbs, dense) if d and torch.is_tensor(Vb)]
                ctx.n_vb = len(saved)
                ctx.save_for_backward(Xb, vbs, gp.get_vs().detach(), *saved)
            else:
                nll = gp._coefficients(X, list(Vs), False)["nll"]
        return nll

    @staticmethod
    def backward(ctx, g):
        Xb, vbs, vs, *Vbs = ctx.saved_tensors
        gmin, gmax = torch.aminmax(g.detach())
        gbar = float(gmax)
        if float(gmin) != gbar:
            raise NotImplementedError(
                "GP.nll: backward is implemented for a uniform upstream gradient only (nll.sum(), nll.mean(), or a "
                "constant multiple); a per-row weighting couples the rows through K^-1 and needs a second solve")
        gX = gbar * Xb if ctx.needs_input_grad[1] else None
        glvs = gbar * vs * (vbs - (vbs * vs).sum()) if ctx.needs_input_grad[2] else None
        it = iter(Vbs)
        gVs = []
        for d, need in zip(ctx.dense, ctx.needs_input_grad[3:]):
            Vb = next(it) if (d and ctx.n_vb) else None
            gVs.append(gbar * Vb if (need and Vb is not None) else None)
        return (None, gX, glvs, *gVs)
