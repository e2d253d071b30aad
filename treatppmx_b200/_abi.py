"""ctypes mirror of include/tppmx.h (the C ABI).  Pure declarations + host-side array
holders; no compute.  Shared by the product wrapper (engine.py) and by the test-only oracle
binding (tests/oracle_binding.py), because both speak the same structs.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

MAXD, MAXQ, MAXCC, MAXT = 8, 16, 8, 16

# RNG sites (include/tppmx.h)
SITE = dict(
    INIT_SS=1, INIT_ETA=2, INIT_EMPTY=3, AUX_SWEEP=4, AUX_PATIENT=5, ALLOC_U=6, AUX_REFRESH=7,
    ETA_PROP=8, HIER_W=9, HIER_THETA=10, SIGKAP=11, BETA=12, HS_LAMBDA=13, HS_TAU=14, JJ=15,
    SS=16, ISPRED=17, PRED_U=18, PRED_ETA=19, PRED_Y=20,
)

ERR_CAPACITY = 3

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class Model(C.Structure):
    _fields_ = [
        ("PPMx", C.c_int32), ("cohesion", C.c_int32), ("similarity", C.c_int32),
        ("consim", C.c_int32), ("calibration", C.c_int32), ("coardegree", C.c_int32),
        ("reuse", C.c_int32), ("CC", C.c_int32), ("upd_hier", C.c_int32), ("hsp", C.c_int32),
        ("noprog", C.c_int32), ("nolik", C.c_int32),
        ("alpha", C.c_double), ("similparam", C.c_double * 7),
        ("hP0_mu0", C.c_double * MAXD),
        ("hP0_nu0", C.c_double), ("hP0_s0", C.c_double), ("hP0_Lambda0", C.c_double),
        ("mhtune_eta", C.c_double), ("mhtune_beta", C.c_double),
    ]


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "nobs", "T", "D", "Q", "P", "ncat", "maxC", "npred", "G", "ntmax", "kcap", "reserved0")]


class Shared(C.Structure):
    _fields_ = [("catvec", c_ip), ("grid", c_dp), ("logV", c_dp)]


class Dataset(C.Structure):
    _fields_ = [("treat", c_ip), ("y", c_dp), ("z", c_dp), ("xcon", c_dp), ("xcat", c_ip),
                ("zpred", c_dp), ("xconp", c_dp), ("xcatp", c_ip)]


_STATE_FIELDS = [
    ("labels", c_ip), ("nj", c_ip), ("nclu", c_ip), ("sumx", c_dp), ("sumx2", c_dp),
    ("njc", c_ip), ("eta_star", c_dp), ("eta_empty", c_dp), ("JJ", c_dp), ("ss", c_dp),
    ("loggamma", c_dp), ("beta", c_dp), ("sigma", c_dp), ("kappa", c_dp), ("idxs", c_ip),
    ("Theta", c_dp), ("Sigma", c_dp), ("lambda_", c_dp), ("tau", c_dp), ("sigma_beta", c_dp),
]


class State(C.Structure):
    _fields_ = _STATE_FIELDS


class OracleProblem(C.Structure):
    _fields_ = [("model", Model), ("dims", Dims), ("shared", Shared), ("data", Dataset)]


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else c_dp()


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else c_ip()


def default_model(**kw) -> Model:
    """Defaults of ppmxct() (reference R/dm_ppmx_ct.R:89-93).  similparam / modelpriors have
    no defaults in the reference; the values below are the ones in its source comment
    (src/dm_ppmx_ct.cpp:222) and are this repo's documented choice."""
    m = Model()
    m.PPMx, m.cohesion, m.similarity, m.consim = 1, 2, 1, 1
    m.calibration, m.coardegree, m.reuse, m.CC = 0, 1, 1, 3
    m.upd_hier, m.hsp, m.noprog, m.nolik = 1, 1, 0, 0
    m.alpha = 1.0
    sp = kw.pop("similparam", (0.0, 10.0, 0.5, 1.0, 2.0, 0.0, 0.1))
    for i, v in enumerate(sp):
        m.similparam[i] = v
    mu0 = kw.pop("hP0_mu0", (0.0,) * MAXD)
    for i, v in enumerate(mu0):
        m.hP0_mu0[i] = v
    m.hP0_nu0, m.hP0_s0, m.hP0_Lambda0 = 10.0, 5.0, 10.0
    m.mhtune_eta, m.mhtune_beta = 0.05, 0.05
    for k, v in kw.items():
        if not hasattr(m, k):
            raise AttributeError(k)
        setattr(m, k, v)
    return m


def model_as_dict(m: Model) -> dict:
    out = {}
    for name, _ in Model._fields_:
        v = getattr(m, name)
        out[name] = list(v) if hasattr(v, "__len__") else v
    return out


@dataclass
class HostDataset:
    """numpy holders of one dataset in the layouts dm_ppmx_ct receives (:10-23)."""
    treat: np.ndarray      # [nobs] int32, 0-based
    y: np.ndarray          # nobs x D, F-order float64
    z: np.ndarray          # nobs x Q, F-order
    xcon: np.ndarray       # nobs x P, C-order (row-major flatten)
    xcat: np.ndarray       # nobs x ncat int32, C-order
    zpred: np.ndarray      # npred x Q, F-order
    xconp: np.ndarray      # npred x P, C-order
    xcatp: np.ndarray      # npred x ncat int32, C-order

    def __post_init__(self):
        self.treat = np.ascontiguousarray(self.treat, dtype=np.int32)
        self.y = np.asfortranarray(self.y, dtype=np.float64)
        self.z = np.asfortranarray(self.z, dtype=np.float64)
        self.xcon = np.ascontiguousarray(self.xcon, dtype=np.float64)
        self.xcat = np.ascontiguousarray(self.xcat, dtype=np.int32)
        self.zpred = np.asfortranarray(self.zpred, dtype=np.float64)
        self.xconp = np.ascontiguousarray(self.xconp, dtype=np.float64)
        self.xcatp = np.ascontiguousarray(self.xcatp, dtype=np.int32)

    def cstruct(self) -> Dataset:
        """The C view of this dataset (eight pointers into the arrays above); built once."""
        d = self.__dict__.get("_cs")
        if d is None:
            d = Dataset()
            d.treat, d.y, d.z = _ip(self.treat), _dp(self.y), _dp(self.z)
            d.xcon, d.xcat = _dp(self.xcon), _ip(self.xcat)
            d.zpred, d.xconp, d.xcatp = _dp(self.zpred), _dp(self.xconp), _ip(self.xcatp)
            self.__dict__["_cs"] = d
        return d


@dataclass
class HostShared:
    catvec: np.ndarray     # [ncat] int32
    grid: np.ndarray       # 2 x G F-order
    logV: np.ndarray | None  # [T][2][ntmax][G] C-order, or None (cohesion 1)

    def __post_init__(self):
        self.catvec = np.ascontiguousarray(self.catvec, dtype=np.int32)
        self.grid = np.asfortranarray(self.grid, dtype=np.float64)
        if self.logV is not None:
            self.logV = np.ascontiguousarray(self.logV, dtype=np.float64)

    def cstruct(self) -> Shared:
        s = Shared()
        s.catvec, s.grid, s.logV = _ip(self.catvec), _dp(self.grid), _dp(self.logV)
        return s


@dataclass
class HostState:
    """One chain's state in the reference's layouts (include/tppmx.h, tppmx_state)."""
    dims: Dims
    CC: int
    arrays: dict = field(default_factory=dict)

    def __post_init__(self):
        d, CC = self.dims, self.CC
        T, nt, P, D, Q, n = d.T, d.ntmax, max(d.P, 1), d.D, d.Q, d.nobs
        ncc = max(d.ncat * d.maxC, 1)
        i32, f64 = np.int32, np.float64
        spec = dict(
            labels=((T, nt), i32), nj=((T, nt), i32), nclu=((T,), i32),
            sumx=((T, nt * P), f64), sumx2=((T, nt * P), f64), njc=((T, nt * ncc), i32),
            eta_star=((D, nt, T), f64), eta_empty=((D, CC, T), f64),
            JJ=((n, D), f64), ss=((n,), f64), loggamma=((n, D), f64), beta=((Q * D,), f64),
            sigma=((T,), f64), kappa=((T,), f64), idxs=((1,), i32),
            Theta=((D, T), f64), Sigma=((D, D, T), f64),
            lambda_=((Q * D,), f64), tau=((1,), f64), sigma_beta=((Q * D,), f64),
        )
        for k, (shape, dt) in spec.items():
            if k not in self.arrays:
                self.arrays[k] = np.zeros(shape, dtype=dt, order="F")

    def __getattr__(self, name):
        arrays = self.__dict__.get("arrays", {})
        if name in arrays:
            return arrays[name]
        raise AttributeError(name)

    def cstruct(self) -> State:
        s = State()
        for name, typ in _STATE_FIELDS:
            a = self.arrays[name]
            setattr(s, name, a.ctypes.data_as(typ))
        return s

    def copy(self) -> "HostState":
        return HostState(self.dims, self.CC, {k: v.copy(order="F") for k, v in self.arrays.items()})
