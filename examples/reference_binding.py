"""The binding a maintainer of savelovme/adaptive-mcmc would add (e.g. as python/kernels/arwmh_b200.py):
a ctypes stub over the C ABI of include/amcmc.h that takes and returns HOST NumPy arrays, so the
reference's NumPy/JAX-CPU scripts can call the B200 sampler without importing torch.

    from kernels.arwmh_b200 import B200ARWMH
    k = B200ARWMH("eight_schools", y=data["y"], sigma=data["sigma"], num_chains=4096)
    state = k.init(seed=0)
    samples, state = k.run(state, num_steps=550_000, num_warmup=50_000, thinning=50)   # samples: [S, C, d]

Only libamcmc.so + numpy are needed.  (tests/test_gpu_small.py::test_reference_binding_example runs it.)
"""
import ctypes as C
import os

import numpy as np

LIB = os.environ.get("AMCMC_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                               "adaptive_mcmc_b200", "libamcmc.so"))


class _State(C.Structure):  # struct amcmc_state
    _fields_ = [("n_chains", C.c_int64), ("dim", C.c_int32), ("dtype", C.c_int32), ("i", C.c_int64),
                ("z", C.c_void_p), ("potential_energy", C.c_void_p), ("mean_accept_prob", C.c_void_p),
                ("loc", C.c_void_p), ("scale", C.c_void_p), ("log_step_size", C.c_void_p), ("as_change", C.c_void_p)]


class _RunArgs(C.Structure):  # struct amcmc_run_args
    _fields_ = [("n_steps", C.c_int64), ("thinning", C.c_int64), ("collect_start", C.c_int64), ("num_warmup", C.c_int64),
                ("lr_decay", C.c_double), ("target_accept_prob", C.c_double), ("eps", C.c_double),
                ("adapt", C.c_int32), ("rng_mode", C.c_int32), ("seed", C.c_uint64), ("chain_offset", C.c_int64),
                ("normals", C.c_void_p), ("uniforms", C.c_void_p), ("out_z", C.c_void_p),
                ("out_potential_energy", C.c_void_p), ("out_accept", C.c_void_p), ("kernel_kind", C.c_int32), ("impl", C.c_int32)]


_MODELS = {"std_normal": 0, "eight_schools": 1, "kidiq": 2, "diamonds": 3}


class B200ARWMH:
    """Host-array front end: same hyper-parameters as ARWMH.__init__ (python/kernels/arwmh.py:43-45)."""

    def __init__(self, model, num_chains, lr_decay=2 / 3, target_accept_prob=0.234, eps=1e-6, **data):
        self.L = C.CDLL(LIB)
        self.L.amcmc_last_error.restype = C.c_char_p
        if model == "eight_schools":
            arrays, d = [data["y"], data["sigma"]], 10
        elif model == "kidiq":
            arrays, d = [data["kid_score"], data["mom_hs"], data["mom_iq"]], 4
        elif model == "diamonds":
            arrays, d = [np.asarray(data["X"]).ravel(), data["Y"]], np.asarray(data["X"]).shape[1] + 1
        else:
            arrays, d = [], int(data.get("d", 1))
        arrays = [np.ascontiguousarray(a, np.float64) for a in arrays]
        ptrs = (C.POINTER(C.c_double) * max(len(arrays), 1))(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrays])
        lens = (C.c_int64 * max(len(arrays), 1))(*[a.size for a in arrays])
        self.h = C.c_void_p()
        self._ck(self.L.amcmc_model_create(C.byref(self.h), _MODELS[model], 0, d, len(arrays), ptrs, lens))
        self.C, self.d = int(num_chains), d
        self.hp = (lr_decay, target_accept_prob, eps)

    def _ck(self, rc):
        if rc:
            raise (ValueError if rc == -1 else RuntimeError)(self.L.amcmc_last_error().decode())

    def _cstate(self, st):
        s = _State(self.C, self.d, 0, st["i"])
        for f in ("z", "potential_energy", "mean_accept_prob", "loc", "scale", "log_step_size", "as_change"):
            setattr(s, f, st[f].ctypes.data)
        return s

    def init(self, seed=0):
        """ARWMH.init (arwmh.py:84-138) for all chains: q0 ~ U(-2,2), U0, loc = q0, scale = I, ...  Returns host
        arrays in the library's struct-of-arrays layout ([d, C], packed lower triangle [d(d+1)/2, C])."""
        C_, d = self.C, self.d
        f = np.float32
        st = dict(i=0, z=np.zeros((d, C_), f), potential_energy=np.zeros(C_, f), mean_accept_prob=np.zeros(C_, f),
                  loc=np.zeros((d, C_), f), scale=np.zeros((d * (d + 1) // 2, C_), f), log_step_size=np.zeros(C_, f),
                  as_change=np.zeros(C_, f), seed=int(seed))
        cs = self._cstate(st)
        self.L.amcmc_arwmh_init_host.argtypes = [C.c_void_p, C.POINTER(_State), C.c_uint64, C.c_int64, C.c_double, C.c_int]
        self._ck(self.L.amcmc_arwmh_init_host(self.h, C.byref(cs), st["seed"], 0, 2.0, 0))
        return st

    def run(self, st, num_steps, num_warmup=0, thinning=1):
        """K fused ARWMH.sample steps through amcmc_arwmh_run_host (host buffers in, host buffers out)."""
        lr, tgt, eps = self.hp
        S = max(0, (num_steps - num_warmup) // thinning)
        out_z = np.empty((S, self.d, self.C), np.float32)
        out_pe = np.empty((S, self.C), np.float32)
        a = _RunArgs(num_steps, thinning, num_warmup, num_warmup, lr, tgt, eps, 1, 0, st["seed"], 0, None, None,
                     out_z.ctypes.data, out_pe.ctypes.data, None, 0, 0)
        cs = self._cstate(st)
        self._ck(self.L.amcmc_arwmh_run_host(self.h, C.byref(cs), C.byref(a)))
        st["i"] = int(cs.i)
        return np.transpose(out_z, (0, 2, 1)), st


if __name__ == "__main__":
    y = [28, 8, -3, 7, -1, 1, 18, 12]
    sigma = [15, 10, 16, 11, 9, 11, 10, 18]
    k = B200ARWMH("eight_schools", num_chains=4096, y=y, sigma=sigma)
    st = k.init(0)
    z, st = k.run(st, 60000, num_warmup=10000, thinning=50)
    print("samples", z.shape, "posterior mean of mu", float(z[..., 0].mean()), "accept", float(st["mean_accept_prob"].mean()))
