"""Host-side mirror of src/model.jl and src/train.jl: hyper-parameters, derived lengths, the learnable `ucdl`
parameter set and the training loop.  The network itself (forward, reverse pass, AdaBelief) runs in
libmotifs_b200 (csrc/csc.cu); nothing here computes on the CPU.

  Hyperparam, length_info         model.jl:1-37
  ucdl (initialisation)           model.jl:67-106 + randomly_initialize_filters MOTIFs.jl:17-33
  forward_pass_return_loss        model.jl:375-395
  setup_num_epochs, train_ucdl    train.jl:1-57
  code_retrieval                  inference/_1_code_retrieval.jl:33-56
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ._lib import CscModel, Context, Sequences  # noqa: F401
from . import parallel  # noqa: F401


@dataclass
class Hyperparam:                      # model.jl:1-14
    filter_len: int = 8
    M: int = 50
    h: int = 12
    K: int = 24
    q: int = 32
    batch_size: int = 6
    num_pass_xyz: int = 6
    num_pass_df: int = 3
    magnifying_factor: float = 10.0
    gamma: float = 0.1

    @property
    def f_len(self):
        return 4 * self.filter_len

    @property
    def twoM(self):
        return 2 * self.M


@dataclass
class length_info:                     # model.jl:16-37
    L: int
    C: int
    c: int
    l: int
    MB: int
    KB: int
    CS_vlen: int
    last_avail_ind: int

    @staticmethod
    def of(hp: Hyperparam, data):
        L = 4 * data.L
        C = L - hp.f_len + 1
        c = data.L - hp.filter_len + 1
        return length_info(L, C, c, c - hp.h + 1, hp.M * hp.batch_size, hp.K * hp.batch_size, C + L - 1,
                           data.N - data.N % hp.batch_size)


_FIELDS = ["lambda_sparsity", "kappa_sparsity", "lambda_stepsize", "omega_stepsize", "kappa_stepsize", "D", "F", "penalty_xyz", "mu"]


class ucdl:
    """Learnable parameters (model.jl:67-137) as ONE flat Float32 vector in Flux.params order followed by the three
    warm-up scalars; the named views below alias it.  D is Julia (32,1,M) memory order, F is (h,2M,1,K)."""

    def __init__(self, hp: Hyperparam, rng: np.random.Generator | None = None, eta1=np.float32(0.05)):
        self.hp = hp
        rng = rng or np.random.default_rng()
        sizes = {"lambda_sparsity": hp.num_pass_xyz, "kappa_sparsity": hp.num_pass_df, "lambda_stepsize": hp.num_pass_xyz,
                 "omega_stepsize": hp.num_pass_xyz, "kappa_stepsize": hp.num_pass_df, "D": hp.f_len * hp.M,
                 "F": hp.h * hp.twoM * hp.K, "penalty_xyz": hp.num_pass_xyz, "mu": hp.num_pass_df}
        self.offsets, o = {}, 0
        for f in _FIELDS:
            self.offsets[f] = (o, o + sizes[f])
            o += sizes[f]
        self.n_trainable = o
        self.flat = np.zeros(o + 3, np.float32)
        # randomly_initialize_filters: every (position, filter) column is a uniform point of the simplex
        # (spacings of 3 sorted uniforms), then D = sqrt.(D)   (MOTIFs.jl:17-33, model.jl:85-89)
        u = np.sort(rng.random((hp.M, hp.filter_len, 3)), axis=2)
        edges = np.concatenate([np.zeros((hp.M, hp.filter_len, 1)), u, np.ones((hp.M, hp.filter_len, 1))], axis=2)
        self["D"][:] = np.sqrt(np.diff(edges, axis=2)).astype(np.float32).reshape(-1)          # [m][j][a] = k + 32 m
        self["F"][:] = np.abs(np.float32(0.1) * rng.standard_normal(sizes["F"]).astype(np.float32))
        for f in _FIELDS:
            if f not in ("D", "F"):
                self[f][:] = eta1 * rng.random(sizes[f]).astype(np.float32)
        self.flat[o:] = eta1 * rng.random(3).astype(np.float32)        # lambda_sparsity_warmup, lambda_stepsize_warmup, omega_stepsize_warmup

    def __getitem__(self, name):
        a, b = self.offsets[name]
        return self.flat[a:b]

    @property
    def D(self):
        """(32, 1, M) Julia order -> numpy (M, 32) C-order view."""
        return self["D"].reshape(self.hp.M, self.hp.f_len)

    @property
    def F(self):
        """(h, 2M, 1, K) Julia order -> numpy (K, 2M, h) C-order view."""
        return self["F"].reshape(self.hp.K, self.hp.twoM, self.hp.h)


def prep_syntax_filters(F):
    """model.jl:148-151 on the numpy (K, 2M, h) view."""
    F2 = np.asarray(F, np.float32) ** 2
    return F2 / np.sqrt((F2 ** 2).sum(axis=(1, 2), keepdims=True))


def forward_pass_return_loss(model: CscModel, seqs: Sequences, batch_idx):
    """loss of one batch (model.jl:375-395) -> float; gradients are available through model.loss_grad."""
    loss, _ = model.loss_grad(seqs, batch_idx, want_grads=False)
    return float(loss[:, 0].mean())


def setup_num_epochs(number_training_samples):   # train.jl:1-11
    if number_training_samples < 1000:
        return 25
    if number_training_samples < 10000:
        return 10
    if number_training_samples < 100000:
        return 5
    return 3


def train_ucdl(data, num_epochs=None, l1_loss_thresh=np.float32(95.0), rng: np.random.Generator | None = None,
               cdl: ucdl | None = None, verbose=True, max_steps=None, groups_per_rank: int = 1, on_step=None):
    """train_ucdl (train.jl:13-57): DataLoader(batch 6, shuffle, partial=false) -> loss + gradient -> AdaBelief ->
    early stop on l1(F) < 95.  Returns (cdl, hp, len, projs=None, model).

    Data parallel (one process per GPU, communicator on the ctx: parallel.init_comm): rank 0's parameters and its epoch
    permutations are broadcast, every rank takes its own batch(es) of 6 from each global step, and the gradients are averaged
    with ONE all-reduce per step inside the library.  The early-stop statistic is computed from bit-identical
    parameters in a fixed order, so all ranks leave the loop on the same step.  With one rank and groups_per_rank=1 this is
    exactly the reference's step sequence."""
    hp = Hyperparam()
    ln = length_info.of(hp, data)
    rng = rng or np.random.default_rng()
    cdl = cdl or ucdl(hp, rng)
    seqs = data.seqs
    ctx = seqs.ctx
    rank, world, _ = ctx.comm_info()
    model = CscModel(ctx, hp, data.L, n_groups=groups_per_rank)
    model.set_params(cdl.flat)
    if world > 1:
        # nothing guarantees that the callers seeded their generators alike: rank 0's weights and shuffles are everybody's
        model.broadcast_params(0)
        seed = np.array([rng.integers(0, 2 ** 62)], np.int64)
        ctx.comm_broadcast(seed, 0)
        rng = np.random.default_rng(int(seed[0]))
    num_epochs = setup_num_epochs(data.N) if num_epochs is None else num_epochs
    per_step = hp.batch_size * groups_per_rank * world
    steps_per_epoch = data.N // per_step
    step, stop = 0, False
    for epoch in range(1, num_epochs + 1):
        perm = rng.permutation(data.N)                                   # shuffle=true; identical on every rank (broadcast seed)
        for it in range(steps_per_epoch):
            lo = it * per_step + rank * hp.batch_size * groups_per_rank
            idx = perm[lo: lo + hp.batch_size * groups_per_rank]
            model.step_begin(seqs, idx)
            loss, l1 = model.adabelief_step()                            # averages the 30 433 gradients over ranks first (one all-reduce)
            step += 1
            if on_step is not None:
                on_step(step, loss, l1)
            if verbose:
                print(f"loss {loss}")                                    # model.jl:392
            if l1 < l1_loss_thresh or (max_steps is not None and step >= max_steps):
                stop = True
                break
        if stop:
            break
        if verbose:
            print(f"Epoch: {epoch} completed")                           # train.jl:55
    cdl.flat[:] = model.get_params()
    return cdl, hp, ln, None, model


class _DevArray:
    """minimal __cuda_array_interface__ wrapper so torch can view library-owned device memory without a copy."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def code_retrieval(data, cdl: ucdl, hp: Hyperparam, model: CscModel | None = None, groups_per_call: int = 1024, tensor_cores: bool = False):
    """code_retrieval (_1_code_retrieval.jl:33-56) -> structured array (position, fil, seq, mag_f16), 0-based, ordered by
    seq, fil, position.  groups_per_call batches of 6 are decoded per kernel sequence (they are independent); with a
    communicator on the ctx the batches are sharded over the ranks and the records all-gathered."""
    seqs = data.seqs
    world = seqs.ctx.world
    n_groups = max(1, min(groups_per_call, -(-(data.N // hp.batch_size) // world)))
    # tensor_cores=True routes the dense syntax-filter contraction through the tcgen05/TMEM BF16 kernel (stated tolerance, not the
    # fp32 parity path)
    m = CscModel(seqs.ctx, hp, data.L, n_groups=n_groups, forward_only=True, tensor_cores=tensor_cores)
    m.set_params(cdl.flat)
    out = m.codes(seqs, shard="comm" if world > 1 else None)
    m.free()
    return out
