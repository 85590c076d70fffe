"""``B200GPSurrogate`` -- drop-in for proFit's ``GPSurrogate`` ("Custom", profit/sur/gp/custom_surrogate.py:7-282).

Same method set, argument meaning, attribute names and error behaviour; the numerics run in libgpb on the GPU.
When ``profit`` is importable the class subclasses ``profit.sur.gp.GaussianProcess`` and registers itself as
``Surrogate["B200"]`` (config: ``fit: {surrogate: B200}`` + ``include: <repo>/profit_b200_include.py``); call
``register_as_custom()`` to also take over the hard-coded ``Surrogate["Custom"]`` of McmcAL (mcmc_al.py:101-102).
Without ``profit`` (e.g. on a bare GPU box) a minimal stand-in base class restates the checks of
``Surrogate.pre_train`` / ``pre_predict`` (sur.py:136-164,187-221) and ``GaussianProcess.pre_train``
(gaussian_process.py:74-136) so the class works stand-alone.
"""
import numpy as np

from . import gp_functions as _gpf
from .backend import GPBackend
from .kernels import kernel_kind, select_kernel

try:  # pragma: no cover - depends on the environment
    from profit.sur import Surrogate as _Surrogate
    from profit.sur.gp import GaussianProcess as _GPBase
    from profit.defaults import fit_gaussian_process as _gp_defaults, fit as _base_defaults

    HAVE_PROFIT = True
except Exception:  # profit not installed: stand-alone base with the same contract
    HAVE_PROFIT = False
    _Surrogate = None
    _gp_defaults = {"surrogate": "GPyTorch", "kernel": "RBF",
                    "hyperparameters": {"length_scale": None, "sigma_f": None, "sigma_n": None}}
    _base_defaults = {"fixed_sigma_n": False}

    class _GPBase:  # restatement of Surrogate + GaussianProcess plumbing (no numerics)
        labels = {}

        def __init__(self):
            self.trained = False
            self.fixed_sigma_n = False
            self.Xtrain = None
            self.ytrain = None
            self.ndim = None
            self.output_ndim = 1
            self.input_encoders = []
            self.output_encoders = []
            self.kernel = None
            self.hyperparameters = {}

        def _check_training_data(self, X, y):            # sur.py:136-164
            X, y = np.asarray(X), np.asarray(y)
            if X.ndim == 1:
                X = X.reshape(-1, 1)
            elif X.ndim != 2:
                raise ValueError(f"X should have shape (n,) or (n, d) but has shape {X.shape}")
            if y.ndim == 1:
                y = y.reshape(-1, 1)
            elif y.ndim != 2:
                raise ValueError(f"y should have shape (n,) or (n, D) but has shape {y.shape}")
            if X.shape[0] != y.shape[0]:
                raise ValueError(f"mismatched number of data points for X and y: {X.shape[0]} != {y.shape[0]}")
            self.Xtrain, self.ytrain = X, y
            self.ndim = self.Xtrain.shape[-1]

        def pre_train(self, X, y, kernel=_gp_defaults["kernel"], hyperparameters=_gp_defaults["hyperparameters"],
                      fixed_sigma_n=_base_defaults["fixed_sigma_n"]):     # gaussian_process.py:74-114
            self._check_training_data(X, y)
            self.set_attributes(fixed_sigma_n=fixed_sigma_n, kernel=kernel, hyperparameters=hyperparameters)
            inferred = None
            for key, value in self.hyperparameters.items():
                if value is None:
                    if inferred is None:
                        inferred = self.infer_hyperparameters()
                    value = inferred[key]
                self.hyperparameters[key] = np.atleast_1d(value)
            if isinstance(self.kernel, str):
                self.kernel = self.select_kernel(self.kernel)

        def set_attributes(self, **kwargs):              # gaussian_process.py:116-122
            for key, value in kwargs.items():
                if not getattr(self, key):
                    new_value = _gp_defaults.get(key, value) if not value else value
                    if isinstance(new_value, dict):
                        new_value = new_value.copy()
                    setattr(self, key, new_value)

        def pre_predict(self, Xpred):                    # sur.py:187-221
            if not self.trained:
                raise RuntimeError("Need to train() before predict()!")
            if Xpred is None:                            # sur.py:200-201
                Xpred = self.default_Xpred()
            Xpred = np.asarray(Xpred)
            if Xpred.ndim == 1:
                Xpred = Xpred.reshape(-1, 1)
            elif Xpred.ndim != 2:
                if self.ndim == 1:
                    raise ValueError(f"Xpred should have shape (n,) or (n, 1) but has shape {Xpred.shape}")
                raise ValueError(f"Xpred should have shape (n, {self.ndim}) but has shape {Xpred.shape}")
            if Xpred.shape[1] != self.ndim:
                raise ValueError(f"Xpred should have shape (n, {self.ndim}) but has shape {Xpred.shape}")
            return Xpred

        def default_Xpred(self):                         # sur.py:399-420
            if self.ndim > 3:
                raise RuntimeError("Require x for prediction in > 3 dimensions!")
            lo, hi = self.Xtrain.min(axis=0), self.Xtrain.max(axis=0)
            axes = [np.linspace(a, b, 50) for a, b in zip(lo, hi)]
            return np.hstack([g.flatten().reshape(-1, 1) for g in np.meshgrid(*axes)])

        def decode_training_data(self):
            pass

        def encode_training_data(self):
            pass

        def encode_predict_data(self, x):
            return x

        def decode_predict_data(self, ym, yv):
            return ym, yv

        def decode_hyperparameters(self):                # gaussian_process.py:222-233 (no encoders stand-alone)
            for key, value in self.hyperparameters.items():
                self.hyperparameters[key] = self.special_hyperparameter_decoding(key, value)

        def special_hyperparameter_decoding(self, key, value):
            return np.atleast_1d(value)

        def post_train(self):                            # sur.py:166-168
            self.trained = True
            self.decode_training_data()

        def print_hyperparameters(self, prefix):
            pass

        @classmethod
        def from_config(cls, config, base_config):       # gaussian_process.py:193-208
            self = cls()
            self.kernel = config["kernel"]
            self.hyperparameters = config["hyperparameters"]
            return self


class B200GPSurrogate(_GPBase):
    """Custom GP on the B200: train / optimize / predict / add_training_data / set_ytrain, reference semantics.

    Attributes (as read by proFit's callers and tests): ``trained, fixed_sigma_n, Xtrain, ytrain, ndim,
    output_ndim, kernel, hyperparameters`` (dict of 1-D float64 arrays), ``hess_inv``.
    """

    gp_functions = _gpf

    def __init__(self, device=None):
        super().__init__()
        self.hess_inv = None
        self.non_spd_evaluations = 0
        self._device = device
        self._backend = None      # may be shared between surrogates (multi-output children): residency is tracked on it

    # -- device plumbing --------------------------------------------------------------------------------
    @property
    def backend(self):
        if self._backend is None:
            self._backend = GPBackend(self._device)
        return self._backend

    @staticmethod
    def _data_key(X, y):
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        return (X.shape, y.shape, hash(X.tobytes()), hash(y.tobytes())) if X.size < (1 << 22) else None

    def _upload(self):
        X = np.ascontiguousarray(self.Xtrain, dtype=np.float64)
        y = np.ascontiguousarray(self.ytrain, dtype=np.float64)
        key = self._data_key(X, y)
        be = self.backend
        if key is None or key != be.resident_key:
            be.set_train(X, y)
            be.resident_key = key

    def _factor_key(self, resident_key):
        return (resident_key, self._theta().tobytes(), getattr(self.kernel, "__name__", str(self.kernel)))

    def _theta(self, with_sigma_n=True):
        keys = ("length_scale", "sigma_f", "sigma_n")
        parts = [np.atleast_1d(np.asarray(self.hyperparameters[k], dtype=np.float64)).ravel() for k in keys]
        return np.concatenate(parts)

    # -- reference API ----------------------------------------------------------------------------------
    def infer_hyperparameters(self):
        """gaussian_process.py:124-136, with the pairwise mean distance evaluated on the device (same value, no
        (N, N, D) tensor)."""
        std = np.std(self.ytrain, axis=0)
        spread = np.abs(self.ytrain.max(axis=0) - self.ytrain.min(axis=0))
        dist = self.backend.mean_pairwise_distance(np.ascontiguousarray(self.Xtrain, dtype=np.float64))
        std, spread = np.mean(std), np.mean(spread)
        return {"length_scale": np.array([0.5 * dist]), "sigma_f": np.array([std]),
                "sigma_n": 1e-2 * np.array([spread])}

    def pre_train(self, X, y, kernel=_gp_defaults["kernel"], hyperparameters=_gp_defaults["hyperparameters"],
                  fixed_sigma_n=_base_defaults["fixed_sigma_n"]):
        if HAVE_PROFIT:
            # Same as GaussianProcess.pre_train (gaussian_process.py:74-114) except that hyper-parameters are
            # only inferred when one is missing: the reference always builds an (N, N, D) tensor here.
            _Surrogate.pre_train(self, X, y)
            self.set_attributes(fixed_sigma_n=fixed_sigma_n, kernel=kernel, hyperparameters=hyperparameters)
            inferred = None
            for key, value in self.hyperparameters.items():
                if value is None:
                    if inferred is None:
                        inferred = self.infer_hyperparameters()
                    value = inferred[key]
                self.hyperparameters[key] = np.atleast_1d(value)
            self.print_hyperparameters("Initial")
            if isinstance(self.kernel, str):
                self.kernel = self.select_kernel(self.kernel)
        else:
            super().pre_train(X, y, kernel, hyperparameters, fixed_sigma_n)

    def train(self, X, y, kernel=_gp_defaults["kernel"], hyperparameters=_gp_defaults["hyperparameters"],
              fixed_sigma_n=_base_defaults["fixed_sigma_n"], eval_gradient=True, return_hess_inv=False):
        """custom_surrogate.py:47-90."""
        self.pre_train(X, y, kernel, hyperparameters, fixed_sigma_n)
        if self.input_encoders or self.output_encoders:
            print("For now, encoding is not supported for this surrogate. The model is created without encoding.")
            self.decode_training_data()
        self.kernel = self.select_kernel(self.kernel)
        self.optimize(fixed_sigma_n=self.fixed_sigma_n, eval_gradient=eval_gradient, return_hess_inv=return_hess_inv)
        self.post_train()

    def post_train(self):
        self.trained = True

    def predict(self, Xpred, add_data_variance=True):
        """custom_surrogate.py:95-120: returns (M, n_out) mean and (M, 1) variance."""
        Xpred = self.pre_predict(Xpred)
        for enc in self.input_encoders[::-1]:
            Xpred = enc.decode(Xpred)
        self._ensure_factor()
        jittered = self.backend.noise_override is not None
        mean, var = self.backend.predict(np.ascontiguousarray(Xpred, dtype=np.float64),
                                         add_data_variance=1 if (add_data_variance and not jittered) else 0)
        if add_data_variance and jittered:
            var += self.backend.noise_override   # the reported data variance stays the model's sigma_n^2
        return mean, var.reshape(-1, 1)

    def predict_into(self, Xpred, out_mean=None, out_var=None, add_data_variance=True):
        """``predict`` for callers that keep everything on the device: ``Xpred`` (M, d), ``out_mean`` (M, n_out) and
        ``out_var`` (M,) are torch CUDA tensors (or numpy arrays); either output may be None to skip it."""
        self._ensure_factor()
        jittered = self.backend.noise_override is not None
        self.backend.predict(Xpred, add_data_variance=1 if (add_data_variance and not jittered) else 0,
                             out_mean=out_mean, out_var=out_var, want_mean=out_mean is not None,
                             want_var=out_var is not None)
        if add_data_variance and jittered and out_var is not None:
            out_var += self.backend.noise_override

    def _ensure_factor(self):
        """Factor of the current (data, hyper-parameters) on the device.  Like the reference's predict path, a
        numerically non-SPD Ky is retried once with 1e-6 on the diagonal (gp_functions.invert, :343-346; alpha falls
        back to lstsq there, custom_surrogate.py:42-45); the data variance reported by predict stays sigma_n^2."""
        self._upload()
        be = self.backend
        key = self._factor_key(be.resident_key)
        if be.resident_key is None or key != be.factor_key:
            theta = self._theta()
            try:
                be.factorize(kernel_kind(self.kernel), theta)
            except np.linalg.LinAlgError:
                jit = np.array(theta)
                jit[-1] = np.sqrt(theta[-1] ** 2 + 1e-6)
                be.factorize(kernel_kind(self.kernel), jit)
                be.noise_override = float(theta[-1] ** 2)     # sigma_n^2 to report for this (jittered) factor
            be.factor_key = key

    def add_training_data(self, X, y, extend_factor=True):
        """custom_surrogate.py:122-134: append rows; hyper-parameters are not re-optimised here.

        The reference leaves "Update Ky" as a TODO (:132) and refactorises on the next predict.  Here, when the
        device holds the factor of exactly this model, the factor is extended in place (O(N^2) per row) -- unless the
        caller says ``extend_factor=False`` because a re-optimisation follows at once (the acquisition loop,
        aquisition_functions.py:70-72), which would discard the extended factor anyway."""
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        be = self._backend
        extend = False
        if extend_factor and be is not None and self.trained and be.resident_key is not None and callable(self.kernel):
            old_key = self._data_key(self.Xtrain, self.ytrain)
            extend = old_key == be.resident_key and be.factor_key == self._factor_key(old_key)
        self.Xtrain = np.concatenate([self.Xtrain, X], axis=0)
        self.ytrain = np.concatenate([self.ytrain, y], axis=0)
        if extend:
            try:
                be.append_train(X.reshape(-1, self.Xtrain.shape[1]), y.reshape(X.shape[0], -1))
            except np.linalg.LinAlgError:
                return      # left to the next predict, which refactorises (with the jitter retry of _ensure_factor)
            be.resident_key = self._data_key(self.Xtrain, self.ytrain)
            be.factor_key = self._factor_key(be.resident_key)

    def set_ytrain(self, ydata):
        """custom_surrogate.py:136-142."""
        self.ytrain = np.atleast_2d(ydata.copy())

    def get_marginal_variance(self, Xpred):
        """custom_surrogate.py:144-163: re-optimise with the analytic gradient to obtain the inverse Hessian, then
        add the hyper-parameter uncertainty of the mean (BBQ) to the predictive variance."""
        self.optimize(fixed_sigma_n=self.fixed_sigma_n, eval_gradient=True, return_hess_inv=True)
        assert self.hess_inv is not None
        _, fvar = self.predict(Xpred)          # leaves the factor of the optimised hyper-parameters on the device
        # the reference hands the raw Xpred on (custom_surrogate.py:152-163): the same points predict() evaluates
        # after undoing the input encoding, NOT the encoded ones
        Xraw = np.asarray(Xpred, dtype=np.float64)
        if Xraw.ndim == 1:
            Xraw = Xraw.reshape(-1, 1)
        return _gpf.marginal_variance_BBQ(self.Xtrain, self.ytrain, Xraw, self.kernel,
                                          self.hyperparameters, self.hess_inv, self.fixed_sigma_n,
                                          predictive_variance=fvar, backend=self.backend)

    def select_kernel(self, kernel):
        """custom_surrogate.py:213-238: name or callable -> kernel callable; unknown names raise RuntimeError."""
        return select_kernel(kernel)

    def optimize(self, fixed_sigma_n=_base_defaults["fixed_sigma_n"], eval_gradient=False, return_hess_inv=False):
        """custom_surrogate.py:240-272."""
        ordered = ("length_scale", "sigma_f", "sigma_n")
        a0 = np.concatenate([np.atleast_1d(self.hyperparameters[key]) for key in ordered])
        self.kernel = self.select_kernel(self.kernel)
        fixed = bool(self.fixed_sigma_n or fixed_sigma_n)
        opt = _gpf.optimize(self.Xtrain, self.ytrain, a0, self.kernel, fixed_sigma_n=fixed,
                            eval_gradient=eval_gradient, return_hess_inv=return_hess_inv, backend=self.backend,
                            resident_key=self._data_key(self.Xtrain, self.ytrain))
        if return_hess_inv:
            self.hess_inv = opt[1]
            opt = opt[0]
        # how many objective evaluations of this fit hit a non-SPD covariance matrix (gp_functions.NonSPDWarning)
        self.non_spd_evaluations = len(self.backend.non_spd_events)
        self._set_hyperparameters_from_model(opt, fixed)
        self.print_hyperparameters("Optimized")

    def _set_hyperparameters_from_model(self, model_hyperparameters, fixed=None):
        # custom_surrogate.py:274-282 (the reference keys on self.fixed_sigma_n only; an optimize(fixed_sigma_n=True)
        # call on an unfixed surrogate would otherwise mis-slice, so the effective flag is used)
        fixed = self.fixed_sigma_n if fixed is None else fixed
        last_idx = -1 if fixed else -2
        self.hyperparameters["length_scale"] = np.atleast_1d(model_hyperparameters[:last_idx])
        self.hyperparameters["sigma_f"] = np.atleast_1d(model_hyperparameters[last_idx])
        if not fixed:
            self.hyperparameters["sigma_n"] = np.atleast_1d(model_hyperparameters[-1])

    def save_model(self, path):
        """custom_surrogate.py:165-187: same dictionary of attributes; .npz here (h5py-free), .hdf5 via proFit's
        FileHandler when it is available."""
        saved_attrs = ("trained", "fixed_sigma_n", "Xtrain", "ytrain", "ndim", "kernel", "hyperparameters")
        save_dict = {attr: getattr(self, attr) for attr in saved_attrs}
        if not isinstance(save_dict["kernel"], str):
            save_dict["kernel"] = self.kernel.__name__
        if str(path).endswith(".npz"):
            flat = {k: v for k, v in save_dict.items() if k != "hyperparameters"}
            for k, v in save_dict["hyperparameters"].items():
                flat["hyperparameters." + k] = v
            np.savez(path, **flat)
        else:
            from profit.util.file_handler import FileHandler

            FileHandler.save(path, save_dict)

    @classmethod
    def load_model(cls, path):
        """custom_surrogate.py:189-211."""
        self = cls()
        if str(path).endswith(".npz"):
            data = np.load(path, allow_pickle=False)
            hyp = {}
            for k in data.files:
                if k.startswith("hyperparameters."):
                    hyp[k.split(".", 1)[1]] = data[k]
                elif k in ("trained", "fixed_sigma_n"):
                    setattr(self, k, bool(data[k]))
                elif k == "ndim":
                    setattr(self, k, int(data[k]))
                elif k == "kernel":
                    setattr(self, k, str(data[k]))
                else:
                    setattr(self, k, data[k])
            self.hyperparameters = hyp
        else:
            from profit.util.file_handler import FileHandler

            for attr, value in FileHandler.load(path, as_type="dict").items():
                setattr(self, attr, value)
        self.kernel = self.select_kernel(self.kernel)
        self.print_hyperparameters("Loaded")
        return self


class B200MultiOutputGPSurrogate(_GPBase):
    """Independent GP per output column -- mirror of ``MultiOutputGPSurrogate`` ("CustomMultiOutputGP",
    custom_surrogate.py:285-451) with ``B200GPSurrogate`` children, including its encoder handling: the parent
    encodes the training data (``Surrogate.pre_train``), the children are trained on the ENCODED arrays and carry no
    encoders of their own, hyper-parameters and predictions are decoded by the parent (:292-350).

    Every child owns a device handle, so its training set and factor stay resident (a prediction over all outputs
    does not refactorise) and the children are fitted ``concurrent`` at a time on host threads -- a single
    evaluation cannot fill the GPU while its Cholesky walks the panel chain, independent fits overlap
    (SURVEY.md section 8 f-3).  ``share_backend=True`` keeps one handle for all children (least memory: each
    switch between outputs then re-uploads y and refactorises)."""

    def __init__(self, child=B200GPSurrogate, device=None, concurrent=2, share_backend=False):
        super().__init__()
        self.child = child
        self.models = []
        self._device = device
        self.concurrent = int(concurrent)
        self.share_backend = bool(share_backend)
        self._shared_backend = None

    def _new_child(self):
        m = self.child(self._device) if self.child is B200GPSurrogate else self.child()
        if self.share_backend or self._shared_backend is not None:
            if self._shared_backend is None:
                self._shared_backend = GPBackend(self._device)
            m._backend = self._shared_backend
        return m

    def infer_hyperparameters(self):
        """gaussian_process.py:124-136 on the (encoded) training data, pairwise mean distance on the device."""
        std = np.mean(np.std(self.ytrain, axis=0))
        spread = np.mean(np.abs(self.ytrain.max(axis=0) - self.ytrain.min(axis=0)))
        be = self._shared_backend or GPBackend(self._device)
        dist = be.mean_pairwise_distance(np.ascontiguousarray(self.Xtrain, dtype=np.float64))
        if be is not self._shared_backend:
            be.close()
        return {"length_scale": np.array([0.5 * dist]), "sigma_f": np.array([std]), "sigma_n": 1e-2 * np.array([spread])}

    def pre_train(self, X, y, kernel=_gp_defaults["kernel"], hyperparameters=_gp_defaults["hyperparameters"],
                  fixed_sigma_n=_base_defaults["fixed_sigma_n"]):
        """GaussianProcess.pre_train (gaussian_process.py:74-114): check + ENCODE the training data, take kernel /
        hyper-parameters / fixed_sigma_n from the config, the arguments or the defaults, infer what is missing.  Only
        difference: the inference runs when a value is actually missing (the reference always builds an (N, N, D)
        tensor on the host)."""
        if HAVE_PROFIT:
            _Surrogate.pre_train(self, X, y)          # sur.py:136-164, encodes
        else:
            self._check_training_data(X, y)
        self.output_ndim = self.ytrain.shape[-1]
        self.set_attributes(fixed_sigma_n=fixed_sigma_n, kernel=kernel, hyperparameters=hyperparameters)
        inferred = None
        for key, value in self.hyperparameters.items():
            if value is None:
                if inferred is None:
                    inferred = self.infer_hyperparameters()
                value = inferred[key]
            self.hyperparameters[key] = np.atleast_1d(value)
        self.print_hyperparameters("Initial")
        if isinstance(self.kernel, str):
            self.kernel = self.select_kernel(self.kernel)

    def train(self, X, y, kernel=_gp_defaults["kernel"], hyperparameters=_gp_defaults["hyperparameters"],
              fixed_sigma_n=_base_defaults["fixed_sigma_n"], return_hess_inv=False):
        """custom_surrogate.py:292-315.  Like the reference, the children are trained with eval_gradient=False
        (the positional ``False`` at :309), i.e. SciPy finite-differences the device NLL, on the encoded data; then the
        hyper-parameters are merged and decoded and ``post_train`` decodes the parent's training arrays."""
        self.pre_train(X, y, kernel, hyperparameters, fixed_sigma_n)
        self.models = [self._new_child() for _ in range(self.output_ndim)]

        def fit(dim):
            self.models[dim].train(self.Xtrain, self.ytrain[:, [dim]], self.kernel, dict(self.hyperparameters),
                                   self.fixed_sigma_n, False, return_hess_inv)

        workers = 1 if self._shared_backend is not None else max(1, min(self.concurrent, self.output_ndim))
        if workers == 1:
            for dim in range(self.output_ndim):
                fit(dim)
        else:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(max_workers=workers) as pool:     # the C calls release the GIL
                list(pool.map(fit, range(self.output_ndim)))
        self._set_hyperparameters_from_model()
        self.post_train()                                             # sur.py:166-168: trained, decode training data

    def _set_hyperparameters_from_model(self):
        # custom_surrogate.py:317-330 verbatim in effect: the children's values are appended to the parent's own
        # (initial) ones, then reduced to norm / max, then decoded
        for m in self.models:
            for k, v in m.hyperparameters.items():
                self.hyperparameters[k] = np.concatenate([np.atleast_1d(self.hyperparameters[k]), np.atleast_1d(v)])
        self.hyperparameters["length_scale"] = np.atleast_1d(np.linalg.norm(self.hyperparameters["length_scale"]))
        self.hyperparameters["sigma_f"] = np.atleast_1d(np.max(self.hyperparameters["sigma_f"]))
        self.hyperparameters["sigma_n"] = np.atleast_1d(np.max(self.hyperparameters["sigma_n"]))
        self.decode_hyperparameters()

    def predict(self, Xpred, add_data_variance=True):
        """custom_surrogate.py:332-340: encode Xpred (pre_predict), children on encoded points, decode the outputs."""
        Xpred = self.pre_predict(Xpred)
        ypred = np.empty((Xpred.shape[0], self.output_ndim))
        yvar = np.empty_like(ypred)
        for dim, m in enumerate(self.models):
            ypred[:, [dim]], yvar[:, [dim]] = m.predict(Xpred, add_data_variance)
        return self.decode_predict_data(ypred, yvar)

    def add_training_data(self, X, y):
        """custom_surrogate.py:342-350: append raw rows, encode everything, hand the encoded new rows to the children,
        decode the parent's copy again."""
        start = self.Xtrain.shape[0]
        self.Xtrain = np.concatenate([self.Xtrain, X], axis=0)
        self.ytrain = np.concatenate([self.ytrain, y], axis=0)
        self.encode_training_data()
        for dim, m in enumerate(self.models):
            m.add_training_data(self.Xtrain[start:], self.ytrain[start:, [dim]])
        self.decode_training_data()

    def set_ytrain(self, y):
        for dim, m in enumerate(self.models):
            m.set_ytrain(y[:, [dim]])
        self.ytrain = np.atleast_2d(y.copy())

    def optimize(self, **opt_kwargs):
        for m in self.models:
            m.optimize(**opt_kwargs)

    def select_kernel(self, kernel):
        return select_kernel(kernel)

    def special_hyperparameter_decoding(self, key, value):
        """custom_surrogate.py:437-444."""
        if len(value) > 1:
            return np.atleast_1d(np.linalg.norm(value)) if key == "length_scale" else np.atleast_1d(np.max(value))
        return value

    def get_marginal_variance(self, Xpred):
        """custom_surrogate.py:446-451: sum of the children's marginal variances at the encoded points."""
        Xpred = self.encode_predict_data(Xpred)
        v = np.zeros((Xpred.shape[0], 1))
        for m in self.models:
            v += m.get_marginal_variance(Xpred)
        return v

    _PARENT_ATTRS = ("Xtrain", "ytrain", "trained", "output_ndim", "hyperparameters")
    _CHILD_ATTRS = ("trained", "fixed_sigma_n", "Xtrain", "ytrain", "ndim", "output_ndim", "kernel", "hyperparameters")

    def _save_dict(self):
        """The dictionary custom_surrogate.py:357-390 hands to FileHandler.save."""
        save_dict = {attr: getattr(self, attr) for attr in self._PARENT_ATTRS}
        save_dict["input_encoders"] = str([enc.repr for enc in self.input_encoders])
        save_dict["output_encoders"] = str([enc.repr for enc in self.output_encoders])
        for i, m in enumerate(self.models):
            save_dict[i] = {attr: getattr(m, attr) for attr in self._CHILD_ATTRS}
            if not isinstance(save_dict[i]["kernel"], str):
                save_dict[i]["kernel"] = m.kernel.__name__
        return save_dict

    def save_model(self, path):
        """custom_surrogate.py:352-390: the same nested dictionary (parent arrays, encoder reprs, one group per child).
        ``.hdf5`` goes through proFit's FileHandler (needs h5py); ``.npz`` is the h5py-free container of the same keys
        ("<child>/<attr>", "hyperparameters/<name>")."""
        save_dict = self._save_dict()
        if str(path).endswith(".npz"):
            flat = {}
            for k, v in save_dict.items():
                if isinstance(v, dict):
                    for kk, vv in v.items():
                        if isinstance(vv, dict):
                            for k3, v3 in vv.items():
                                flat[f"{k}/{kk}/{k3}"] = v3
                        else:
                            flat[f"{k}/{kk}"] = vv
                else:
                    flat[str(k)] = v
            np.savez(path, **flat)
            return
        from profit.util.file_handler import FileHandler

        FileHandler.save(path, save_dict)

    @classmethod
    def load_model(cls, path):
        """custom_surrogate.py:392-431: restore the parent, rebuild its encoders (and initialise them by one encode /
        decode round trip), then the children."""
        if str(path).endswith(".npz"):
            data = np.load(path, allow_pickle=False)
            load_dict = {}
            for key in data.files:
                parts = key.split("/")
                node = load_dict
                for part in parts[:-1]:
                    node = node.setdefault(part, {})
                val = data[key]
                node[parts[-1]] = val.item() if val.ndim == 0 else val
        else:
            from profit.util.file_handler import FileHandler

            load_dict = FileHandler.load(path, as_type="dict")
        self = cls()
        self.trained = bool(load_dict["trained"])
        self.output_ndim = int(load_dict["output_ndim"])
        self.Xtrain, self.ytrain = load_dict["Xtrain"], load_dict["ytrain"]
        self.hyperparameters = dict(load_dict["hyperparameters"])
        self.ndim = self.Xtrain.shape[-1]
        if HAVE_PROFIT:
            from numpy import array  # noqa: F401  (needed by eval of the encoder reprs, as upstream)
            from profit.sur.encoders import Encoder

            for enc in eval(str(load_dict["input_encoders"])):
                self.add_input_encoder(Encoder[enc["class"]](enc["columns"], enc["parameters"]))
            for enc in eval(str(load_dict["output_encoders"])):
                self.add_output_encoder(Encoder[enc["class"]](enc["columns"], enc["parameters"]))
            self.encode_training_data()
            self.decode_training_data()
        self.models = [self._new_child() for _ in range(self.output_ndim)]
        for i, m in enumerate(self.models):
            for attr, value in load_dict[str(i)].items():
                if attr in ("trained", "fixed_sigma_n"):
                    value = bool(value)
                elif attr in ("ndim", "output_ndim"):
                    value = int(value)
                elif attr == "kernel":
                    value = str(value)
                elif attr == "hyperparameters":
                    value = dict(value)
                setattr(m, attr, value)
            m.kernel = m.select_kernel(m.kernel)
            m.print_hyperparameters("Loaded")
        return self


def register(label="B200"):
    """Register the class in proFit's surrogate registry (util/base_class.py:15-26)."""
    if not HAVE_PROFIT:
        raise RuntimeError("profit is not importable: nothing to register with")
    return _Surrogate.register(label)(B200GPSurrogate)


def register_as_custom():
    """Take over the label "Custom" (prints proFit's duplicate-label warning, base_class.py:20-23)."""
    return register("Custom")


if HAVE_PROFIT:  # pragma: no cover
    try:
        register("B200")
        _Surrogate.register("B200MultiOutputGP")(B200MultiOutputGPSurrogate)
    except Exception:
        pass
