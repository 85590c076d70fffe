"""Acquisition functions of proFit's active learning on the device.

Mirror of ``profit/al/aquisition_functions.py``: same class names, constructor arguments, ``calculate_loss`` /
``find_next_candidates`` contract and registry labels (prefixed ``b200_``; ``register_as_default()`` takes over the
reference's own labels).  The candidate set is uploaded once; every pick runs ``gpb_predict`` into device buffers
and ``gpb_acquire`` (loss + ``np.argmax(loss * mask)``) on the device, so 16 bytes come back per pick instead of
2·M doubles, and ``add_training_data`` extends the factor in place (``gpb_append_train``).

The surrogate must be a ``B200GPSurrogate`` (it provides ``predict_into`` and the device backend); there is no
CPU path.  ``torch`` only supplies the device buffers.

Under ``torch.distributed`` (one process per GPU, every rank running the same active-learning loop on the same
training data) the candidate set is sharded automatically: each rank predicts and scores its contiguous shard, the
min / max that the normalisations need are combined with one small all_gather, and the pick is one all_gather of
(value, global index) -- so every rank arrives at the same point and keeps an identical model.  ``af.sharded =
False`` switches that off.
"""
import numpy as np

from .backend import ACQ_DISTANCE_PENALTY, ACQ_EI, ACQ_EI2, ACQ_POI, ACQ_VARIANCE, ACQ_WEIGHTED
from .distributed import gather_argmax, gather_concat, gather_minmax, shard_bounds, world

try:  # pragma: no cover - depends on the environment
    from profit.al.aquisition_functions import AcquisitionFunction as _RefBase

    HAVE_PROFIT = True
except Exception:
    HAVE_PROFIT = False

    class _RefBase:  # registry + constructor of the reference base class (aquisition_functions.py:26-56)
        labels = {}
        al_parameters = {}
        EPSILON = 1e-12

        def __init__(self, Xpred, surrogate, variables, **parameters):
            self.parameters = parameters
            self.Xpred = Xpred
            self.surrogate = surrogate
            self.variables = variables

        def set_al_parameters(self, **kwargs):
            for key, value in kwargs.items():
                if key in self.al_parameters:
                    self.al_parameters[key] = value
                else:
                    print(f"Skipped setting AL parameter {key}.")

        @classmethod
        def register(cls, label):
            def decorator(obj):
                if label in cls.labels:
                    print(f"registering duplicate label '{label}' for {cls.__name__}.")
                cls.labels[label] = obj
                return obj

            return decorator

        def __class_getitem__(cls, item):
            return cls if item is None else cls.labels[item]


# defaults of profit/defaults.py:93-126
_DEFAULTS = {
    "use_marginal_variance": False,
    "distance_weight": 10,
    "exploration_weight": 0.5,
    "exploration_factor": 0.01,
    "find_min": False,
    "alternating_freq": 1,
}


class DeviceAcquisitionFunction(_RefBase):
    """Base: device buffers, the fused predict -> loss -> masked arg-max pick, and the batch loop."""

    kind = None
    sharded = None          # None: shard the candidates when torch.distributed runs more than one rank

    def __init__(self, Xpred, surrogate, variables, **parameters):
        super().__init__(Xpred, surrogate, variables, **parameters)
        self._dev = None

    # -- device plumbing ------------------------------------------------------------------------------
    def _shard(self, M):
        rank, ws = world()
        if self.sharded is False or ws == 1:
            return 0, M, False
        lo, hi = shard_bounds(M, ws, rank)
        return lo, hi, True

    def _buffers(self):
        if not hasattr(self.surrogate, "predict_into"):
            raise TypeError("profit_b200 acquisition functions need a B200GPSurrogate (device predict); "
                            f"got {type(self.surrogate).__name__}")
        import torch

        Xp = np.ascontiguousarray(self.Xpred, dtype=np.float64)
        if Xp.ndim == 1:
            Xp = Xp.reshape(-1, 1)
        d = self._dev
        n_out = int(np.atleast_2d(self.surrogate.ytrain).shape[1])
        if d is None or d["src"] is not self.Xpred or d["n_out"] != n_out:
            dev = torch.device("cuda", self.surrogate.backend.device)
            lo, hi, sharded = self._shard(Xp.shape[0])
            M = hi - lo                                   # candidates held by this rank (all of them if not sharded)
            d = {"src": self.Xpred, "n_out": n_out, "M": M, "M_total": Xp.shape[0], "lo": lo, "hi": hi,
                 "sharded": sharded,
                 "X": torch.from_numpy(Xp[lo:hi]).to(dev),
                 "mean": torch.empty((M, n_out), dtype=torch.float64, device=dev),
                 "var": torch.empty(M, dtype=torch.float64, device=dev),
                 "term": torch.empty((M, n_out), dtype=torch.float64, device=dev)}
            self._dev = d
        return d

    def _predict(self, want_mean=False, want_var=True):
        """Predictive mean / variance of this rank's candidates into the device buffers (nothing returns to the host)."""
        d = self._buffers()
        if d["M"] > 0:
            self.surrogate.predict_into(d["X"], out_mean=d["mean"] if want_mean else None,
                                        out_var=d["var"] if want_var else None, add_data_variance=True)
        return d

    def _variance(self, use_marginal_variance=False):
        """Device predictive variance, or the host marginal variance (aquisition_functions.py:108-118)."""
        if use_marginal_variance:
            if hasattr(self.surrogate, "get_marginal_variance"):
                d = self._buffers()
                mv = np.ascontiguousarray(self.surrogate.get_marginal_variance(self.Xpred), dtype=np.float64).ravel()
                return np.ascontiguousarray(mv[d["lo"]:d["hi"]])
            print("Surrogate has no method 'get_marginal_variance'. Using predictive variance instead.")
        return self._predict()["var"]

    def _prepare(self, kind, params=()):
        """gpb_acq_prepare of the device mean into the ``term`` buffer, with the min / max of the normalisation
        taken over all ranks' shards."""
        d = self._predict(want_mean=True, want_var=False)
        be = self.surrogate.backend
        if not d["sharded"]:
            be.acq_prepare(kind, d["mean"], params=params, out=d["term"])
            return d["term"]
        local = np.tile([np.inf, -np.inf], (d["n_out"], 1))
        if d["M"] > 0:
            _, local = be.acq_prepare(kind, d["mean"], params=params, out=d["term"], return_stats=True)
        stats = gather_minmax(local)
        if d["M"] > 0:
            be.acq_prepare(kind, d["mean"], params=params, out=d["term"], stats_in=stats)
        return d["term"]

    def _inputs(self, *loss_args):
        """-> dict(kind, var, term, extra, params) for gpb_acquire (arrays cover this rank's candidates)."""
        raise NotImplementedError

    def _rows(self, arr):
        """This rank's rows of a per-candidate argument: device tensors from ``normalized_mean`` / ``mu_part`` are
        already local; a host array over all candidates (what the reference's callers pass) is sliced."""
        d = self._buffers()
        if d["sharded"] and isinstance(arr, np.ndarray) and arr.shape[0] == d["M_total"]:
            return np.ascontiguousarray(arr[d["lo"]:d["hi"]])
        return arr

    _NEEDS_STATS = {ACQ_DISTANCE_PENALTY: 0, ACQ_WEIGHTED: 0, ACQ_EI: 1}     # kind -> transform of the variance

    def _acquire(self, visited, loss_out, *loss_args):
        """(global index, value) of ``np.argmax(loss * mask)``; ``loss_out`` (M_local,) receives this rank's losses."""
        a = self._inputs(*loss_args)
        d = self._buffers()
        be = self.surrogate.backend
        stats = None
        if d["sharded"] and a["kind"] in self._NEEDS_STATS:
            local = be.acq_stats(a["var"], self._NEEDS_STATS[a["kind"]]) if d["M"] > 0 else [[np.inf, -np.inf]]
            stats = gather_minmax(local)[0]
        idx, val = -1, -np.inf
        if d["M"] > 0:
            mine = [v - d["lo"] for v in visited if d["lo"] <= v < d["hi"]]
            idx, val = be.acquire(a["kind"], a["var"], term=a.get("term"), extra=a.get("extra"),
                                  params=a.get("params", ()), visited=mine, n_out=d["n_out"], loss_out=loss_out,
                                  stats=stats)
            idx += d["lo"]
        return gather_argmax(val, idx) if d["sharded"] else (idx, val)

    # -- reference API --------------------------------------------------------------------------------
    def calculate_loss(self, *loss_args):
        """Loss per candidate as a host array (M,), like the reference's ``calculate_loss``."""
        d = self._buffers()
        loss = np.empty(d["M"])
        self._acquire((), loss if d["M"] > 0 else None, *loss_args)
        return gather_concat(loss, d["M_total"]) if d["sharded"] else loss

    def find_next_candidates(self, batch_size):
        return self._find_next_candidates(batch_size)

    def _find_next_candidates(self, batch_size, *loss_args):
        """aquisition_functions.py:63-74 with the pick on the device."""
        Xp = np.asarray(self.Xpred)
        candidates = np.empty((batch_size, Xp.shape[-1]))
        visited = []
        for n in range(batch_size):
            idx, _ = self._acquire(visited, None, *loss_args)
            visited.append(idx)                      # mask[idx] = 0
            candidates[n] = Xp[idx]
            mu_candidate, _ = self.surrogate.predict(candidates[n].reshape(1, -1))
            # B200GPSurrogate extends the device-resident factor and training set in place (O(N^2)), so the optimize()
            # that follows finds the data resident and uploads nothing (gp_functions.optimize, resident_key)
            self.surrogate.add_training_data(candidates[n].reshape(1, -1), mu_candidate)
            self.surrogate.optimize()
        return candidates

    def _observed_max(self, nan_aware):
        out = np.atleast_2d(np.asarray(self.variables.output, dtype=np.float64))
        return np.atleast_1d(np.nanmax(out, axis=0) if nan_aware else out.max(axis=0))


@_RefBase.register("b200_simple_exploration")
class SimpleExploration(DeviceAcquisitionFunction):
    """Next points where the (predictive or marginal) variance is largest (aquisition_functions.py:87-120)."""

    kind = ACQ_VARIANCE

    def __init__(self, Xpred, surrogate, variables, use_marginal_variance=_DEFAULTS["use_marginal_variance"],
                 **parameters):
        super().__init__(Xpred, surrogate, variables, use_marginal_variance=use_marginal_variance, **parameters)

    def _inputs(self):
        return {"kind": ACQ_VARIANCE, "var": self._variance(self.parameters["use_marginal_variance"])}


@_RefBase.register("b200_exploration_with_distance_penalty")
class ExplorationWithDistancePenalty(SimpleExploration):
    """Variance plus ``max(var) * (1 - exp(-weight * |x - x_last|))``, halved (aquisition_functions.py:123-158)."""

    kind = ACQ_DISTANCE_PENALTY

    def __init__(self, Xpred, surrogate, variables, use_marginal_variance=_DEFAULTS["use_marginal_variance"],
                 weight=_DEFAULTS["distance_weight"]):
        super().__init__(Xpred, surrogate, variables, use_marginal_variance=use_marginal_variance, weight=weight)

    def _inputs(self):
        vin = np.asarray(self.variables.input, dtype=np.float64)
        last_point = np.atleast_1d(vin[np.sum(~np.isnan(vin), axis=0).min() - 1])        # :150-152
        d = self._buffers()
        params = np.concatenate([[float(self.parameters["weight"]), float(last_point.size)], last_point])
        return {"kind": ACQ_DISTANCE_PENALTY, "var": self._variance(self.parameters["use_marginal_variance"]),
                "extra": d["X"], "params": params}


@_RefBase.register("b200_weighted_exploration")
class WeightedExploration(DeviceAcquisitionFunction):
    """``weight * normalize(mu) + (1 - weight) * normalize(var)`` (aquisition_functions.py:161-201)."""

    kind = ACQ_WEIGHTED

    def __init__(self, Xpred, surrogate, variables, weight=_DEFAULTS["exploration_weight"],
                 use_marginal_variance=_DEFAULTS["use_marginal_variance"]):
        super().__init__(Xpred, surrogate, variables, weight=weight, use_marginal_variance=use_marginal_variance)

    def _inputs(self, mu):
        return {"kind": ACQ_WEIGHTED, "var": self._variance(self.parameters["use_marginal_variance"]),
                "term": self._rows(mu), "params": [float(self.parameters["weight"])]}

    def normalized_mean(self):
        """``normalize(mu)`` of the current model, kept on the device (:199-200)."""
        return self._prepare(ACQ_WEIGHTED)

    def find_next_candidates(self, batch_size):
        return self._find_next_candidates(batch_size, self.normalized_mean())


@_RefBase.register("b200_probability_of_improvement")
class ProbabilityOfImprovement(DeviceAcquisitionFunction):
    """One-sample improvement ``max(mu + sigma * eps - y_max, 0)`` (aquisition_functions.py:204-224).  The draws
    come from numpy's global generator exactly like the reference's, so a seeded run picks the same points."""

    kind = ACQ_POI

    def _inputs(self, mu):
        d = self._predict()
        noise = np.random.standard_normal((d["M_total"], d["n_out"]))                    # :215, same draws on every rank
        noise = np.ascontiguousarray(noise[d["lo"]:d["hi"]])
        return {"kind": ACQ_POI, "var": d["var"], "term": self._rows(mu), "extra": noise,
                "params": self._observed_max(nan_aware=False)}                            # :216 (plain max, as written)

    def predicted_mean(self):
        d = self._predict(want_mean=True, want_var=False)
        d["term"].copy_(d["mean"])
        return d["term"]

    def find_next_candidates(self, batch_size):
        return self._find_next_candidates(batch_size, self.predicted_mean())


@_RefBase.register("b200_expected_improvement")
class ExpectedImprovement(DeviceAcquisitionFunction):
    """Expected improvement with the mean part frozen over a batch (aquisition_functions.py:227-280)."""

    kind = ACQ_EI
    SIGMA_EPSILON = 1e-10

    def __init__(self, Xpred, surrogate, variables, exploration_factor=_DEFAULTS["exploration_factor"],
                 find_min=_DEFAULTS["find_min"]):
        super().__init__(Xpred, surrogate, variables, exploration_factor=exploration_factor, find_min=find_min)

    def _inputs(self, improvement):
        return {"kind": ACQ_EI, "var": self._predict()["var"], "term": self._rows(improvement)}

    def mu_part(self):
        """``normalize(max(+-mu - nanmax(y), 0)) - xi`` on the device (:262-269)."""
        params = np.concatenate([[float(self.parameters["exploration_factor"]), float(bool(self.parameters["find_min"]))],
                                 self._observed_max(nan_aware=True)])
        return self._prepare(ACQ_EI, params)

    def find_next_candidates(self, batch_size):
        return self._find_next_candidates(batch_size, self.mu_part())


@_RefBase.register("b200_expected_improvement_2")
class ExpectedImprovement2(DeviceAcquisitionFunction):
    """First point by expected improvement, the rest of the batch by largest variance
    (aquisition_functions.py:283-339).  No re-optimisation inside the batch, so every added point is an in-place
    extension of the device factor.  (The reference passes the ``(mean, var)`` tuple of ``predict`` to
    ``add_training_data`` at :323-324/:332-335, which raises in numpy; the mean is used here.)"""

    kind = ACQ_EI2

    def __init__(self, Xpred, surrogate, variables, exploration_factor=_DEFAULTS["exploration_factor"],
                 find_min=_DEFAULTS["find_min"]):
        super().__init__(Xpred, surrogate, variables, exploration_factor=exploration_factor, find_min=find_min)

    def _inputs(self):
        d = self._predict(want_mean=True)
        params = np.concatenate([[float(self.parameters["exploration_factor"]), float(bool(self.parameters["find_min"]))],
                                 self._observed_max(nan_aware=True)])
        return {"kind": ACQ_EI2, "var": d["var"], "term": d["mean"], "params": params}

    def find_next_candidates(self, batch_size):
        Xp = np.asarray(self.Xpred)
        candidates = np.empty((batch_size, Xp.shape[-1]))
        idx, _ = self._acquire((), None)
        candidates[0] = Xp[idx]
        mu_candidate, _ = self.surrogate.predict(candidates[0].reshape(1, -1))
        self.surrogate.add_training_data(candidates[0].reshape(1, -1), mu_candidate)
        exploration = SimpleExploration(self.Xpred, self.surrogate, self.variables)
        exploration._dev = self._dev                       # share the resident candidate set
        for n in range(1, batch_size):
            idx, _ = exploration._acquire((), None)
            candidates[n] = Xp[idx]
            mu_candidate, _ = self.surrogate.predict(candidates[n].reshape(1, -1))
            self.surrogate.add_training_data(candidates[n].reshape(1, -1), mu_candidate)
        return candidates


@_RefBase.register("b200_alternating_exploration")
class AlternatingAF(DeviceAcquisitionFunction):
    """Alternates between variance and expected improvement every ``alternating_freq`` runs
    (aquisition_functions.py:342-380)."""

    al_parameters = {"krun": 0}

    def __init__(self, Xpred, surrogate, variables, use_marginal_variance=_DEFAULTS["use_marginal_variance"],
                 exploration_factor=_DEFAULTS["exploration_factor"], find_min=_DEFAULTS["find_min"],
                 alternating_freq=_DEFAULTS["alternating_freq"]):
        super().__init__(Xpred, surrogate, variables, alternating_freq=alternating_freq)
        self.exploration = SimpleExploration(Xpred, surrogate, variables, use_marginal_variance=use_marginal_variance)
        self.expected_improvement = ExpectedImprovement(Xpred, surrogate, variables,
                                                        exploration_factor=exploration_factor, find_min=find_min)
        self.current_af = self.exploration

    def find_next_candidates(self, batch_size):
        if not self.al_parameters["krun"] % self.parameters["alternating_freq"]:
            self.current_af = (self.expected_improvement if self.current_af is self.exploration
                               else self.exploration)
        return self.current_af.find_next_candidates(batch_size)


_TWINS = {
    "simple_exploration": SimpleExploration,
    "exploration_with_distance_penalty": ExplorationWithDistancePenalty,
    "weighted_exploration": WeightedExploration,
    "probability_of_improvement": ProbabilityOfImprovement,
    "expected_improvement": ExpectedImprovement,
    "expected_improvement_2": ExpectedImprovement2,
    "alternating_exploration": AlternatingAF,
}


def register_as_default():
    """Take over the reference's own labels (``acquisition_function: {class: simple_exploration}`` then selects the
    device twin); prints proFit's duplicate-label notice (util/base_class.py:20-23)."""
    for label, cls in _TWINS.items():
        _RefBase.register(label)(cls)
