"""Sharding the hot path over the GPUs of one box.

The path shards along independent filters, so there is no collective inside a
filter: every rank (one process per GPU, ``torch.distributed``) runs whole filters
on its own device and only log-likelihood vectors are gathered (NCCL between GPUs,
gloo in the CPU tests).

  * pMCMC chains        -> ``mcstate_b200.pmcmc.pmcmc_sharded`` (one chain per GPU)
  * nested populations  -> ``nested_filter_sharded`` here: the populations of ONE
    multi-parameter model are split into contiguous blocks, one block per rank.  The
    unsharded model owns one RNG stream per particle and one shared filter stream
    (tests/testthat/test-particle-filter.R:1316-1318 pins the stream partition); each
    shard takes the streams of its own particles and replays the shared filter
    stream's whole draw order (``mcb_set_stream_sharing``), so the gathered result is
    bit-identical to the unsharded filter's, whatever the number of ranks.
  * SMC^2 parameter particles -> ``mcstate_b200.smc2`` (per-set seeds, so placement
    never matters).
"""
from __future__ import annotations

import numpy as np

from .dust import dust_rng
from .particle_filter import particle_filter
from .particle_filter_data import subset_populations


def block_range(n_items, rank, world):
    """Contiguous block [lo, hi) of ``n_items`` for ``rank``: sizes differ by at most one,
    the larger blocks first (so two ranks of three populations hold 2 + 1)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("sharded filters need an initialised torch.distributed process "
                           "group (one rank per GPU)")
    return dist


def gather_log_likelihood(local, counts, group=None, device=None):
    """All-gather the per-rank log-likelihood vectors (8 bytes x populations): the only
    traffic of a sharded filter evaluation.  ``counts[r]`` = entries rank r contributes."""
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    width = max(counts)
    buf = torch.full((width,), float("nan"), dtype=torch.float64, device=device)
    buf[:len(local)] = torch.as_tensor(np.asarray(local, dtype=np.float64), device=device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)])


class likelihood_gather:
    """The all-gather of per-rank log-likelihood vectors with everything allocated once: a
    device staging vector, the gathered device matrix and a pinned host copy.  With NCCL the
    local vector can be the handle's own device memory (``gather_device``), so a data row of
    SMC^2 costs one stream synchronisation: filter -> all-gather -> device-to-host copy are
    queued on one stream and waited for together (R/smc2.R:89-108 gathers one step
    likelihood per parameter particle per row)."""

    def __init__(self, counts, group=None):
        import torch

        self._dist = _dist()
        self._group = group
        self.counts = [int(c) for c in counts]
        self.world = len(self.counts)
        self.rank = self._dist.get_rank(group)
        self.width = max(self.counts)
        self.nccl = self._dist.get_backend(group) == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if self.nccl else None
        self._local = torch.full((self.width,), float("nan"), dtype=torch.float64, device=dev)
        self._all = torch.empty((self.world * self.width,), dtype=torch.float64, device=dev)
        self._host = torch.empty((self.world * self.width,), dtype=torch.float64,
                                 pin_memory=self.nccl)
        self._take = np.concatenate([np.arange(c) + r * self.width
                                     for r, c in enumerate(self.counts)])

    def _finish(self, wait):
        import torch

        self._dist.all_gather_into_tensor(self._all, self._local, group=self._group)
        if self.nccl:
            self._host.copy_(self._all, non_blocking=True)
            if wait:
                torch.cuda.current_stream().synchronize()
        else:
            self._host.copy_(self._all)

    def gather_host(self, local):
        """local: this rank's values on the host -> every rank's, concatenated."""
        import torch

        n = self.counts[self.rank]
        self._local[:n].copy_(torch.as_tensor(np.asarray(local, dtype=np.float64)),
                              non_blocking=self.nccl)
        self._finish(True)
        return self._host.numpy()[self._take].copy()

    def gather_device(self, local_device, wait=None):
        """local_device: this rank's values as a CUDA tensor written by work already queued
        on the current stream.  ``wait``: a callable that waits for the stream (the handle's
        ``filter_collect``), so the caller synchronises once for the row."""
        n = self.counts[self.rank]
        self._local[:n].copy_(local_device, non_blocking=True)
        self._finish(wait is None)
        if wait is not None:
            wait()
        return self._host.numpy()[self._take].copy()


class nested_filter_sharded:
    """``particle_filter$new(data_nested, model, n_particles, compare = NULL)`` with the
    populations split over the ranks of a process group (BASELINE config C4).

    ``run(pars)`` takes the FULL list of per-population parameters on every rank (as the
    reference's nested filter does, R/particle_filter.R:907-913) and returns the full
    vector of log-likelihoods on every rank.

    Bit-identical to the unsharded filter for every run that ends normally.  When a
    population becomes impossible (all weights -Inf) the RESULT is still the unsharded
    one -- the whole vector -Inf, R/particle_filter_state.R:463-474 -- but only the rank
    that holds the population stops at that row: the other ranks' particle streams run on
    to the end, so the model objects are then in a different state from the unsharded
    filter's (which stopped everywhere).  ``run`` starts every evaluation from
    ``update_state(pars, time)`` and the filter's seed handling, like the reference's
    ``particle_filter$run``; a caller that needs the streams to continue identically after
    an impossible evaluation must recreate the filter.
    """

    def __init__(self, data, model, n_particles, seed=1, index=None, gpu_config=None,
                 group=None):
        dist = _dist()
        if not data.is_nested:
            raise ValueError("nested_filter_sharded needs nested (multi-population) data")
        self._group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        pops = data.populations
        self.n_populations = len(pops)
        if self.n_populations < self.world:
            raise ValueError("more ranks than populations")
        self._lo, self._hi = block_range(self.n_populations, self.rank, self.world)
        self._counts = [block_range(self.n_populations, r, self.world) for r in range(self.world)]
        self._counts = [hi - lo for lo, hi in self._counts]
        self.n_particles = int(n_particles)
        self._gpu_config = gpu_config
        self._tensor_device = None
        if dist.get_backend(group) == "nccl":
            import torch

            self._tensor_device = torch.device("cuda", torch.cuda.current_device())

        # this rank's populations of the data
        local = subset_populations(data, pops[self._lo:self._hi])
        self._filter = particle_filter(local, model, n_particles, None, index=index,
                                       seed=self._local_seed(seed, model), gpu_config=gpu_config)
        # which data rows are NA, for EVERY population: a shard must know when the other
        # populations skip their resampling draw
        field = model.data_fields[0] if hasattr(model, "data_fields") else "incidence"
        vals = np.asarray(data[field], dtype=float)
        pop_col = np.asarray(data[data.population])
        self._na_global = np.stack([np.isnan(vals[pop_col == p]) for p in pops], axis=1)
        self._filter_stream = self._global_filter_stream(seed, model)
        self._attached = False
        self._gather = likelihood_gather(self._counts, group)

    # -- seeds: particle streams of block [lo, hi) start at stream lo * n_particles of the seed
    def _algorithm(self, model):
        return getattr(model, "algorithm", "xoshiro256plus")

    def _local_seed(self, seed, model):
        return dust_rng(seed, self._algorithm(model)).jump(self._lo * self.n_particles).state() \
            if self._lo > 0 else dust_rng(seed, self._algorithm(model)).state()

    def _global_filter_stream(self, seed, model):
        n_total = self.n_populations * self.n_particles
        return dust_rng(seed, self._algorithm(model)).jump(n_total).state()

    def _attach(self):
        # after the first run the model object exists: give it the shared stream and the
        # global draw order (both survive later update_state calls)
        m = self._filter._last_model[0]
        m.set_filter_stream(self._filter_stream)
        m.set_stream_sharing(self.n_populations, self._lo, self._na_global)
        self._attached = True

    def run(self, pars, save_history=False, save_restart=None):
        if isinstance(pars, dict) or len(pars) != self.n_populations:
            raise ValueError("'pars' must be an unnamed list with one entry per population")
        local_pars = list(pars[self._lo:self._hi])
        if not self._attached:
            # create the model object without running the data, then attach the sharing
            obj = self._filter.run_begin(local_pars, save_history, save_restart)
            self._filter._last_model = [obj.model]
            self._attach()
            obj.run()
            self._filter._last_history = obj.history
            self._filter._last_state = obj.model.state
            self._filter._last_restart_state = obj.restart_state
            ll = obj.log_likelihood
        else:
            ll = self._filter.run(local_pars, save_history, save_restart)
        ll = self._gather.gather_host(np.atleast_1d(ll))
        # the reference's rule when any population is impossible: the whole vector
        # (R/particle_filter_state.R:463-474, `ll[] <- -Inf`)
        if np.any(ll == -np.inf):
            ll[:] = -np.inf
        return ll

    # local pieces, for callers that need them (one particle's lineage in pMCMC)
    def state(self, index_state=None):
        return self._filter.state(index_state)

    def history(self, index_particle=None):
        return self._filter.history(index_particle)

    @property
    def populations_local(self):
        return self._lo, self._hi
