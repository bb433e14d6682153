"""Butcher tableaux as runtime data (they travel to the GPU as kernel parameters / constant bank).

Mirrors ``src/solvers/imex_tableaus.jl`` (32 named IMEX tableaux), ``src/solvers/explicit_tableaus.jl``
(SSP22Heuns, SSP33ShuOsher, RK4) and the LSRK coefficient sets of ``src/solvers/lsrk.jl``.  Every
coefficient is evaluated in IEEE double in the expression order of the reference so that the Float64
values are bit-identical (Appendix B of SURVEY.md lists the ARS343 hex values that the tests pin).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from fractions import Fraction as Fr
from typing import Optional, Sequence

import numpy as np

from . import _lib as L

_S2, _S3 = math.sqrt(2.0), math.sqrt(3.0)


def _rowsum(a: np.ndarray) -> np.ndarray:
    """``vec(sum(a; dims = 2))`` with left-to-right accumulation (imex_tableaus.jl:66,69)."""
    out = np.zeros(a.shape[0])
    for i in range(a.shape[0]):
        acc = 0.0
        for j in range(a.shape[1]):
            acc = acc + float(a[i, j])
        out[i] = acc
    return out


@dataclass
class IMEXTableau:
    """imex_tableaus.jl:35-79.  ``b`` defaults to the last row of ``a`` (FSAL), ``c`` to its row sums."""

    a_exp: np.ndarray
    a_imp: np.ndarray
    b_exp: Optional[np.ndarray] = None
    b_imp: Optional[np.ndarray] = None
    c_exp: Optional[np.ndarray] = None
    c_imp: Optional[np.ndarray] = None

    def __post_init__(self):
        f = lambda x: np.array(x, dtype=np.float64)
        self.a_exp, self.a_imp = f(self.a_exp), f(self.a_imp)
        self.b_exp = self.a_exp[-1, :].copy() if self.b_exp is None else f(self.b_exp)
        self.b_imp = self.a_imp[-1, :].copy() if self.b_imp is None else f(self.b_imp)
        self.c_exp = _rowsum(self.a_exp) if self.c_exp is None else f(self.c_exp)
        self.c_imp = _rowsum(self.a_imp) if self.c_imp is None else f(self.c_imp)
        if not np.all(np.triu(self.a_exp) == 0):  # :71
            raise AssertionError("a_exp must be strictly lower triangular")
        if not np.all(np.triu(self.a_imp, 1) == 0):  # :72
            raise AssertionError("a_imp must be lower triangular")
        s = self.a_exp.shape[0]
        if s > L.CTS_MAX_STAGES:
            raise ValueError(f"at most {L.CTS_MAX_STAGES} stages are supported")

    @property
    def s(self) -> int:
        return len(self.b_exp)

    def is_fsal(self) -> bool:  # imex_tableaus.jl:59-62
        return (np.array_equal(self.b_exp, self.a_exp[-1, :]) and np.array_equal(self.b_imp, self.a_imp[-1, :]))

    def to_c(self, cast_to_state_eltype: bool = False) -> L.ImexTableau:
        t = L.ImexTableau()
        s = self.s
        t.s = s
        for i in range(s):
            for j in range(s):
                t.a_exp[i * s + j] = self.a_exp[i, j]
                t.a_imp[i * s + j] = self.a_imp[i, j]
            t.b_exp[i], t.b_imp[i] = self.b_exp[i], self.b_imp[i]
            t.c_exp[i], t.c_imp[i] = self.c_exp[i], self.c_imp[i]
        t.cast_to_state_eltype = int(cast_to_state_eltype)
        return t


def is_fsal(x) -> bool:
    return (x.tableau if hasattr(x, "tableau") else x).is_fsal()


# --------------------------------------------------------------------------------------------------
# algorithm names
# --------------------------------------------------------------------------------------------------
class AbstractAlgorithmName:
    default_constraint = "Unconstrained"  # ClimaTimeSteppers.jl:113

    def tableau(self) -> IMEXTableau:
        raise NotImplementedError

    def __repr__(self):
        return type(self).__name__ + "()"


class IMEXARKAlgorithmName(AbstractAlgorithmName):
    pass


class IMEXSSPRKAlgorithmName(IMEXARKAlgorithmName):
    default_constraint = "SSP"  # imex_tableaus.jl:646


def _name(cls_name, base, fn, doc):
    return type(cls_name, (base,), {"tableau": lambda self: fn(), "__doc__": doc})


def _imkg(alpha, alpha_hat, delta_hat, beta=None) -> IMEXTableau:
    """IMKGTableau (imex_tableaus.jl:382-396)."""
    if beta is None:
        beta = (0,) * len(delta_hat)
    s = len(alpha_hat) + 1
    a_exp, a_imp = np.zeros((s, s)), np.zeros((s, s))
    for i in range(1, s + 1):
        for j in range(1, s + 1):
            a_exp[i - 1, j - 1] = alpha[j - 1] if i == j + 1 else (beta[i - 3] if (i > 2 and j == 1) else 0)
            if i == j + 1:
                v = alpha_hat[j - 1]
            elif i > 2 and j == 1:
                v = beta[i - 3]
            elif 1 < i <= len(alpha_hat) and i == j:
                v = delta_hat[i - 2]
            else:
                v = 0
            a_imp[i - 1, j - 1] = v
    return IMEXTableau(a_exp=a_exp, a_imp=a_imp)


def _ars343():
    g = 0.4358665215084590
    a42 = 0.5529291480359398
    a43 = a42
    b1 = -3 / 2 * g ** 2 + 4 * g - 1 / 4
    b2 = 3 / 2 * g ** 2 - 5 * g + 5 / 4
    a31 = ((1 - 9 / 2 * g + 3 / 2 * g ** 2) * a42 + (11 / 4 - 21 / 2 * g + 15 / 4 * g ** 2) * a43
           - 7 / 2 + 13 * g - 9 / 2 * g ** 2)
    a32 = ((-1 + 9 / 2 * g - 3 / 2 * g ** 2) * a42 + (-11 / 4 + 21 / 2 * g - 15 / 4 * g ** 2) * a43
           + 4 - 25 / 2 * g + 9 / 2 * g ** 2)
    a41 = 1 - a42 - a43
    return IMEXTableau(a_exp=[[0, 0, 0, 0], [g, 0, 0, 0], [a31, a32, 0, 0], [a41, a42, a43, 0]], b_exp=[0, b1, b2, g],
                       a_imp=[[0, 0, 0, 0], [0, g, 0, 0], [0, (1 - g) / 2, g, 0], [0, b1, b2, g]])


def _ars233():
    g = 1 / 2 + _S3 / 6
    return IMEXTableau(a_exp=[[0, 0, 0], [g, 0, 0], [g - 1, 2 - 2 * g, 0]], b_exp=[0, 1 / 2, 1 / 2],
                       a_imp=[[0, 0, 0], [0, g, 0], [0, 1 - 2 * g, g]], b_imp=[0, 1 / 2, 1 / 2])


def _ars232():
    g, d = 1 - _S2 / 2, -2 * _S2 / 3
    return IMEXTableau(a_exp=[[0, 0, 0], [g, 0, 0], [d, 1 - d, 0]], b_exp=[0, 1 - g, g],
                       a_imp=[[0, 0, 0], [0, g, 0], [0, 1 - g, g]])


def _ars222():
    g = 1 - _S2 / 2
    d = 1 - 1 / (2 * g)  # `1 - 1 / 2γ`: juxtaposition binds tighter than `/` in Julia
    return IMEXTableau(a_exp=[[0, 0, 0], [g, 0, 0], [d, 1 - d, 0]], a_imp=[[0, 0, 0], [0, g, 0], [0, 1 - g, g]])


def _ssp333(beta=None):
    beta = 1 / 2 + _S3 / 6 if beta is None else beta
    if not beta > 1 / 2:  # imex_tableaus.jl:733 (test/unit/tableaus.jl:135-139)
        raise AssertionError("SSP333 requires β > 1/2")
    g = (2 * beta ** 2 - 3 * beta / 2 + 1 / 3) / (2 - 4 * beta)
    return IMEXTableau(a_exp=[[0, 0, 0], [1, 0, 0], [1 / 4, 1 / 4, 0]], b_exp=[1 / 6, 1 / 6, 2 / 3],
                       a_imp=[[0, 0, 0], [4 * g + 2 * beta, 1 - 4 * g - 2 * beta, 0], [1 / 2 - beta - g, g, beta]],
                       b_imp=[1 / 6, 1 / 6, 2 / 3])


def _ssp433():
    al, be, et = 0.24169426078821, 0.06042356519705, 0.12915286960590
    return IMEXTableau(a_exp=[[0, 0, 0, 0], [0, 0, 0, 0], [0, 1, 0, 0], [0, 1 / 4, 1 / 4, 0]], b_exp=[0, 1 / 6, 1 / 6, 2 / 3],
                       a_imp=[[al, 0, 0, 0], [-al, al, 0, 0], [0, 1 - al, al, 0], [be, et, 1 / 2 - al - be - et, al]],
                       b_imp=[0, 1 / 6, 1 / 6, 2 / 3])


def _dbm453():
    g = 0.32591194130117247
    return IMEXTableau(
        a_exp=[[0, 0, 0, 0, 0], [0.10306208811591838, 0, 0, 0, 0], [-0.94124866143519894, 1.6626399742527356, 0, 0, 0],
               [-1.3670975201437765, 1.3815852911016873, 1.2673234025619065, 0, 0],
               [-0.81287582068772448, 0.81223739060505738, 0.90644429603699305, 0.094194134045674111, 0]],
        b_exp=[0.87795339639076675, -0.72692641526151547, 0.7520413715737272, -0.22898029400415088, g],
        a_imp=[[0, 0, 0, 0, 0], [-0.2228498531852541, g, 0, 0, 0], [-0.46801347074080545, 0.86349284225716961, g, 0, 0],
               [-0.46509906651927421, 0.81063103116959553, 0.61036726756832357, g, 0],
               [0.87795339639076675, -0.72692641526151547, 0.7520413715737272, -0.22898029400415088, g]])


def _hommem1():
    a_exp = np.zeros((6, 6))
    a_imp = np.zeros((6, 6))
    for i, v in enumerate((1 / 5, 1 / 5, 1 / 3, 1 / 2, 1)):
        a_exp[i + 1, i] = v
    for i, v in enumerate((1 / 5, 1 / 5, 1 / 3, 1 / 2)):
        a_imp[i + 1, i + 1] = v
    a_imp[5, :] = [5 / 18, 5 / 18, 0, 0, 0, 8 / 18]
    return IMEXTableau(a_exp=a_exp, a_imp=a_imp)


def _ark2gkc(paper_version=False):
    a32 = 1 / 2 + _S2 / 3 if paper_version else 1 / 2
    return IMEXTableau(a_exp=[[0, 0, 0], [2 - _S2, 0, 0], [1 - a32, a32, 0]], b_exp=[_S2 / 4, _S2 / 4, 1 - _S2 / 2],
                       a_imp=[[0, 0, 0], [1 - _S2 / 2, 1 - _S2 / 2, 0], [_S2 / 4, _S2 / 4, 1 - _S2 / 2]])


def _rat(table, n):
    m = [[Fr(0)] * n for _ in range(n)]
    for (i, j), (p, q) in table.items():
        m[i - 1][j - 1] = Fr(p, q)
    return m


def _ark437l2sa1():
    """imex_tableaus.jl:880-956 (Rational{Int64} entries converted to Float64 = correctly rounded p/q)."""
    n = 7
    g = Fr(1235, 10000)
    ai = _rat({(3, 2): (624185399699, 4186980696204), (4, 2): (1258591069120, 10082082980243),
               (4, 3): (-322722984531, 8455138723562), (5, 2): (-436103496990, 5971407786587),
               (5, 3): (-2689175662187, 11046760208243), (5, 4): (4431412449334, 12995360898505),
               (6, 2): (-2207373168298, 14430576638973), (6, 3): (242511121179, 3358618340039),
               (6, 4): (3145666661981, 7780404714551), (6, 5): (5882073923981, 14490790706663),
               (7, 3): (9164257142617, 17756377923965), (7, 4): (-10812980402763, 74029279521829),
               (7, 5): (1335994250573, 5691609445217), (7, 6): (2273837961795, 8368240463276)}, n)
    for i in range(1, n):
        ai[i][i] = g
    ae = _rat({(3, 1): (247, 4000), (3, 2): (2694949928731, 7487940209513), (4, 1): (464650059369, 8764239774964),
               (4, 2): (878889893998, 2444806327765), (4, 3): (-952945855348, 12294611323341),
               (5, 1): (476636172619, 8159180917465), (5, 2): (-1271469283451, 7793814740893),
               (5, 3): (-859560642026, 4356155882851), (5, 4): (1723805262919, 4571918432560),
               (6, 1): (6338158500785, 11769362343261), (6, 2): (-4970555480458, 10924838743837),
               (6, 3): (3326578051521, 2647936831840), (6, 4): (-880713585975, 1841400956686),
               (6, 5): (-1428733748635, 8843423958496), (7, 2): (760814592956, 3276306540349),
               (7, 3): (-47223648122716, 6934462133451), (7, 4): (71187472546993, 9669769126921),
               (7, 5): (-13330509492149, 9695768672337), (7, 6): (11565764226357, 8513123442827)}, n)
    b = [Fr(0)] * n
    b[2], b[3] = Fr(9164257142617, 17756377923965), Fr(-10812980402763, 74029279521829)
    b[4], b[5], b[6] = Fr(1335994250573, 5691609445217), Fr(2273837961795, 8368240463276), Fr(247, 2000)
    c = [Fr(0)] * n
    c[1], c[2], c[3], c[4], c[5] = Fr(247, 1000), Fr(4276536705230, 10142255878289), Fr(67, 200), Fr(3, 40), Fr(7, 10)
    for i in range(1, n):
        ai[i][0] = ai[i][1]
    b[0] = b[1]
    ae[1][0] = c[1]
    ae[6][0] = ae[6][1]
    c[0], c[6] = Fr(0), Fr(1)
    fl = lambda m: [[float(x) for x in r] for r in m]
    fv = lambda v: [float(x) for x in v]
    return IMEXTableau(a_exp=fl(ae), b_exp=fv(b), c_exp=fv(c), a_imp=fl(ai), b_imp=fv(b), c_imp=fv(c))


def _ark548l2sa2():
    """imex_tableaus.jl:965-1066."""
    n = 8
    g = Fr(2, 9)
    ai = _rat({(3, 2): (2366667076620, 8822750406821), (4, 2): (-257962897183, 4451812247028),
               (4, 3): (128530224461, 14379561246022), (5, 2): (-486229321650, 11227943450093),
               (5, 3): (-225633144460, 6633558740617), (5, 4): (1741320951451, 6824444397158),
               (6, 2): (621307788657, 4714163060173), (6, 3): (-125196015625, 3866852212004),
               (6, 4): (940440206406, 7593089888465), (6, 5): (961109811699, 6734810228204),
               (7, 2): (2036305566805, 6583108094622), (7, 3): (-3039402635899, 4450598839912),
               (7, 4): (-1829510709469, 31102090912115), (7, 5): (-286320471013, 6931253422520),
               (7, 6): (8651533662697, 9642993110008)}, n)
    for i in range(1, n):
        ai[i][i] = g
    ae = _rat({(3, 1): (1, 9), (3, 2): (1183333538310, 1827251437969), (4, 1): (895379019517, 9750411845327),
               (4, 2): (477606656805, 13473228687314), (4, 3): (-112564739183, 9373365219272),
               (5, 1): (-4458043123994, 13015289567637), (5, 2): (-2500665203865, 9342069639922),
               (5, 3): (983347055801, 8893519644487), (5, 4): (2185051477207, 2551468980502),
               (6, 1): (-167316361917, 17121522574472), (6, 2): (1605541814917, 7619724128744),
               (6, 3): (991021770328, 13052792161721), (6, 4): (2342280609577, 11279663441611),
               (6, 5): (3012424348531, 12792462456678), (7, 1): (6680998715867, 14310383562358),
               (7, 2): (5029118570809, 3897454228471), (7, 3): (2415062538259, 6382199904604),
               (7, 4): (-3924368632305, 6964820224454), (7, 5): (-4331110370267, 15021686902756),
               (7, 6): (-3944303808049, 11994238218192), (8, 1): (2193717860234, 3570523412979),
               (8, 3): (5952760925747, 18750164281544), (8, 4): (-4412967128996, 6196664114337),
               (8, 5): (4151782504231, 36106512998704), (8, 6): (572599549169, 6265429158920),
               (8, 7): (-457874356192, 11306498036315)}, n)
    ae[7][1] = ae[7][0]
    b = [Fr(0)] * n
    b[2], b[3] = Fr(3517720773327, 20256071687669), Fr(4569610470461, 17934693873752)
    b[4], b[5] = Fr(2819471173109, 11655438449929), Fr(3296210113763, 10722700128969)
    b[6] = Fr(-1142099968913, 5710983926999)
    c = [Fr(0)] * n
    c[1], c[2], c[3] = Fr(4, 9), Fr(6456083330201, 8509243623797), Fr(1632083962415, 14158861528103)
    c[4], c[5], c[6] = Fr(6365430648612, 17842476412687), Fr(18, 25), Fr(191, 200)
    for i in range(1, n):
        ai[i][0] = ai[i][1]
    b[0] = b[1]
    b[7] = g
    for i in range(n):
        ai[7][i] = b[i]
    ae[1][0] = c[1]
    ae[7][0] = ae[7][1]
    c[0], c[7] = Fr(0), Fr(1)
    fl = lambda m: [[float(x) for x in r] for r in m]
    fv = lambda v: [float(x) for x in v]
    return IMEXTableau(a_exp=fl(ae), b_exp=fv(b), c_exp=fv(c), a_imp=fl(ai), b_imp=fv(b), c_imp=fv(c))


_A5 = (1 / 4, 1 / 6, 3 / 8, 1 / 2, 1)
_A4 = (1 / 4, 1 / 3, 1 / 2, 1)
_G = 1 - _S2 / 2
_IMEX = {
    # ARS (imex_tableaus.jl:176-353)
    "ARS111": (IMEXARKAlgorithmName, lambda: IMEXTableau(a_exp=[[0, 0], [1, 0]], a_imp=[[0, 0], [0, 1]])),
    "ARS121": (IMEXARKAlgorithmName, lambda: IMEXTableau(a_exp=[[0, 0], [1, 0]], b_exp=[0, 1], a_imp=[[0, 0], [0, 1]])),
    "ARS122": (IMEXARKAlgorithmName, lambda: IMEXTableau(a_exp=[[0, 0], [1 / 2, 0]], b_exp=[0, 1],
                                                        a_imp=[[0, 0], [0, 1 / 2]], b_imp=[0, 1])),
    "ARS233": (IMEXARKAlgorithmName, _ars233),
    "ARS232": (IMEXARKAlgorithmName, _ars232),
    "ARS222": (IMEXARKAlgorithmName, _ars222),
    "ARS343": (IMEXARKAlgorithmName, _ars343),
    "ARS443": (IMEXARKAlgorithmName, lambda: IMEXTableau(
        a_exp=[[0, 0, 0, 0, 0], [1 / 2, 0, 0, 0, 0], [11 / 18, 1 / 18, 0, 0, 0], [5 / 6, -5 / 6, 1 / 2, 0, 0],
               [1 / 4, 7 / 4, 3 / 4, -7 / 4, 0]],
        a_imp=[[0, 0, 0, 0, 0], [0, 1 / 2, 0, 0, 0], [0, 1 / 6, 1 / 2, 0, 0], [0, -1 / 2, 1 / 2, 1 / 2, 0],
               [0, 3 / 2, -3 / 2, 1 / 2, 1 / 2]])),
    # IMKG (imex_tableaus.jl:399-607)
    "IMKG232a": (IMEXARKAlgorithmName, lambda: _imkg((1 / 2, 1 / 2, 1), (0, -1 / 2 + _S2 / 2, 1), (1 - _S2 / 2, 1 - _S2 / 2))),
    "IMKG232b": (IMEXARKAlgorithmName, lambda: _imkg((1 / 2, 1 / 2, 1), (0, -1 / 2 - _S2 / 2, 1), (1 + _S2 / 2, 1 + _S2 / 2))),
    "IMKG242a": (IMEXARKAlgorithmName, lambda: _imkg(_A4, (0, 0, -1 / 2 + _S2 / 2, 1), (0, 1 - _S2 / 2, 1 - _S2 / 2))),
    "IMKG242b": (IMEXARKAlgorithmName, lambda: _imkg(_A4, (0, 0, -1 / 2 - _S2 / 2, 1), (0, 1 + _S2 / 2, 1 + _S2 / 2))),
    "IMKG243a": (IMEXARKAlgorithmName, lambda: _imkg(_A4, (0, 1 / 6, -_S3 / 6, 1),
                                                    (1 / 2 + _S3 / 6, 1 / 2 + _S3 / 6, 1 / 2 + _S3 / 6))),
    "IMKG252a": (IMEXARKAlgorithmName, lambda: _imkg(_A5, (0, 0, 0, -1 / 2 + _S2 / 2, 1), (0, 0, 1 - _S2 / 2, 1 - _S2 / 2))),
    "IMKG252b": (IMEXARKAlgorithmName, lambda: _imkg(_A5, (0, 0, 0, -1 / 2 - _S2 / 2, 1), (0, 0, 1 + _S2 / 2, 1 + _S2 / 2))),
    "IMKG253a": (IMEXARKAlgorithmName, lambda: _imkg(
        _A5, (0, 0, _S3 / 4 * (1 - _S3 / 3) * ((1 + _S3 / 3) ** 2 - 2), _S3 / 6, 1),
        (0, 1 / 2 - _S3 / 6, 1 / 2 - _S3 / 6, 1 / 2 - _S3 / 6))),
    "IMKG253b": (IMEXARKAlgorithmName, lambda: _imkg(
        _A5, (0, 0, _S3 / 4 * (1 + _S3 / 3) * ((1 - _S3 / 3) ** 2 - 2), -_S3 / 6, 1),
        (0, 1 / 2 + _S3 / 6, 1 / 2 + _S3 / 6, 1 / 2 + _S3 / 6))),
    "IMKG254a": (IMEXARKAlgorithmName, lambda: _imkg(_A5, (0, -3 / 10, 5 / 6, -3 / 2, 1), (-1 / 2, 1, 1, 2))),
    "IMKG254b": (IMEXARKAlgorithmName, lambda: _imkg(_A5, (0, -1 / 20, 5 / 4, -1 / 2, 1), (-1 / 2, 1, 1, 1))),
    "IMKG254c": (IMEXARKAlgorithmName, lambda: _imkg(_A5, (0, 1 / 20, 5 / 36, 1 / 3, 1), (1 / 6, 1 / 6, 1 / 6, 1 / 6))),
    "IMKG342a": (IMEXARKAlgorithmName, lambda: _imkg((1 / 4, 2 / 3, 1 / 3, 3 / 4), (0, 1 / 6 - _S3 / 6, -1 / 6 - _S3 / 6, 3 / 4),
                                                    (0, 1 / 2 + _S3 / 6, 1 / 2 + _S3 / 6), (0, 1 / 3, 1 / 4))),
    "IMKG343a": (IMEXARKAlgorithmName, lambda: _imkg((1 / 4, 2 / 3, 1 / 3, 3 / 4), (0, -1 / 3, -2 / 3, 3 / 4), (-1 / 3, 1, 1),
                                                    (0, 1 / 3, 1 / 4))),
    # IMEX SSPRK (imex_tableaus.jl:654-777)
    "SSP222": (IMEXSSPRKAlgorithmName, lambda: IMEXTableau(a_exp=[[0, 0], [1, 0]], b_exp=[1 / 2, 1 / 2],
                                                          a_imp=[[_G, 0], [1 - 2 * _G, _G]], b_imp=[1 / 2, 1 / 2])),
    "SSP322": (IMEXSSPRKAlgorithmName, lambda: IMEXTableau(
        a_exp=[[0, 0, 0], [0, 0, 0], [0, 1, 0]], b_exp=[0, 1 / 2, 1 / 2],
        a_imp=[[1 / 2, 0, 0], [-1 / 2, 1 / 2, 0], [0, 1 / 2, 1 / 2]], b_imp=[0, 1 / 2, 1 / 2])),
    "SSP332": (IMEXSSPRKAlgorithmName, lambda: IMEXTableau(
        a_exp=[[0, 0, 0], [1, 0, 0], [1 / 4, 1 / 4, 0]], b_exp=[1 / 6, 1 / 6, 2 / 3],
        a_imp=[[_G, 0, 0], [1 - 2 * _G, _G, 0], [1 / 2 - _G, 0, _G]], b_imp=[1 / 6, 1 / 6, 2 / 3])),
    "SSP433": (IMEXSSPRKAlgorithmName, _ssp433),
    # miscellaneous (imex_tableaus.jl:789-1066)
    "DBM453": (IMEXARKAlgorithmName, _dbm453),
    "HOMMEM1": (IMEXARKAlgorithmName, _hommem1),
    "ARK437L2SA1": (IMEXARKAlgorithmName, _ark437l2sa1),
    "ARK548L2SA2": (IMEXARKAlgorithmName, _ark548l2sa2),
}
for _n, (_base, _fn) in _IMEX.items():
    globals()[_n] = _name(_n, _base, _fn, f"IMEX tableau {_n} (src/solvers/imex_tableaus.jl)")


class SSP333(IMEXSSPRKAlgorithmName):
    """imex_tableaus.jl:728-748: family parametrised by β (> 1/2)."""

    def __init__(self, beta: Optional[float] = None):
        self.beta = 1 / 2 + _S3 / 6 if beta is None else beta

    def tableau(self):
        return _ssp333(self.beta)


class ARK2GKC(IMEXARKAlgorithmName):
    """imex_tableaus.jl:845-865."""

    def __init__(self, paper_version: bool = False):
        self.paper_version = paper_version

    def tableau(self):
        return _ark2gkc(self.paper_version)


IMEX_NAMES = sorted(list(_IMEX) + ["SSP333", "ARK2GKC"])


# ---- explicit tableaux (explicit_tableaus.jl:82-124) ------------------------------------------
class ERKAlgorithmName(AbstractAlgorithmName):
    def explicit(self):
        raise NotImplementedError

    def tableau(self):
        a, b = self.explicit()
        a = np.array(a, dtype=np.float64)
        b = np.array(b, dtype=np.float64)
        c = _rowsum(a)
        return IMEXTableau(a_exp=a, b_exp=b, c_exp=c, a_imp=a, b_imp=b, c_imp=c)  # IMEXTableau((; a, b, c)) :68


class SSPRKAlgorithmName(ERKAlgorithmName):
    default_constraint = "SSP"


class SSP22Heuns(SSPRKAlgorithmName):
    def explicit(self):
        return [[0, 0], [1, 0]], [1 / 2, 1 / 2]


class SSP33ShuOsher(SSPRKAlgorithmName):
    def explicit(self):
        return [[0, 0, 0], [1, 0, 0], [1 / 4, 1 / 4, 0]], [1 / 6, 1 / 6, 2 / 3]


class RK4(ERKAlgorithmName):
    def explicit(self):
        return [[0, 0, 0, 0], [1 / 2, 0, 0, 0], [0, 1 / 2, 0, 0], [0, 0, 1, 0]], [1 / 6, 1 / 3, 1 / 3, 1 / 6]


EXPLICIT_NAMES = ["SSP22Heuns", "SSP33ShuOsher", "RK4"]


# ---- Rosenbrock (rosenbrock.jl:25-46, 241-267) ---------------------------------------------------
class RosenbrockTableau:
    """``RosenbrockTableau(α, Γ, b)``: stores A = α Γ⁻¹, C = diag(Γ⁻¹) − Γ⁻¹ and m = b Γ⁻¹ next to α and Γ."""

    def __init__(self, alpha, Gamma, b):
        alpha, Gamma, b = (np.array(x, dtype=np.float64) for x in (alpha, Gamma, b))
        n = len(b)
        assert alpha.shape == (n, n) and Gamma.shape == (n, n)
        inv = np.linalg.inv(Gamma)
        if np.allclose(np.diag(Gamma), 1.0) and np.allclose(np.triu(Gamma, 1), 0.0) and n <= 3:
            # unit lower-triangular Γ with dyadic entries: forward substitution is exact
            inv = np.eye(n)
            for i in range(n):
                for j in range(i):
                    inv[i, j] = -sum(Gamma[i, k] * inv[k, j] for k in range(j, i))
        self.alpha, self.Gamma, self.s = alpha, Gamma, n
        self.A = np.array([[sum(alpha[i, k] * inv[k, j] for k in range(n)) for j in range(n)] for i in range(n)])
        self.C = np.diag(np.diag(inv)) - inv
        self.m = np.array([sum(b[k] * inv[k, j] for k in range(n)) for j in range(n)])

    def to_c(self):
        from . import _lib as L
        t = L.RosenbrockTableau()
        t.s = self.s
        for name, M in (("A", self.A), ("alpha", self.alpha), ("C", self.C), ("Gamma", self.Gamma)):
            arr = getattr(t, name)
            for i in range(self.s):
                for j in range(self.s):
                    arr[i * self.s + j] = float(M[i, j])
        for i in range(self.s):
            t.m[i] = float(self.m[i])
        return t


class RosenbrockAlgorithmName(AbstractAlgorithmName):
    pass


class SSPKnoth(RosenbrockAlgorithmName):
    """3-stage, 2nd-order Rosenbrock method (rosenbrock.jl:241-267)."""

    def tableau(self) -> RosenbrockTableau:
        return RosenbrockTableau([[0, 0, 0], [1, 0, 0], [1 / 4, 1 / 4, 0]], [[1, 0, 0], [0, 1, 0], [-3 / 4, -3 / 4, 1]],
                                 [1 / 6, 1 / 6, 2 / 3])


def tableau(name):
    """``CTS.tableau(name)``."""
    return name.tableau()


# ---- low-storage 2N coefficient sets (lsrk.jl:117-217) ----------------------------------------
def _rt(dtype):
    t = np.dtype(dtype).type
    return lambda num, den=None: t(num) if den is None else t(t(num) / t(den))  # RT(num // den)


def lsrk54_carpenter_kennedy(dtype=np.float64):
    q = _rt(dtype)
    A = (q(0), q(-567301805773, 1357537059087), q(-2404267990393, 2016746695238), q(-3550918686646, 2091501179385),
         q(-1275806237668, 842570457699))
    B = (q(1432997174477, 9575080441755), q(5161836677717, 13612068292357), q(1720146321549, 2090206949498),
         q(3134564353537, 4481467310338), q(2277821191437, 14882151754819))
    Cc = (q(0), q(1432997174477, 9575080441755), q(2526269341429, 6820363962896), q(2006345519317, 3224310063776),
          q(2802321613138, 2924317926251))
    return A, B, Cc


def lsrk144_niegemann_diehl_busch(dtype=np.float64):
    q = _rt(dtype)
    A = (0, -0.7188012108672410, -0.7785331173421570, -0.0053282796654044, -0.8552979934029281, -3.9564138245774565,
         -1.5780575380587385, -2.0837094552574054, -0.7483334182761610, -0.7032861106563359, 0.0013917096117681,
         -0.0932075369637460, -0.9514200470875948, -7.1151571693922548)
    B = (0.0367762454319673, 0.3136296607553959, 0.1531848691869027, 0.0030097086818182, 0.3326293790646110,
         0.2440251405350864, 0.3718879239592277, 0.6204126221582444, 0.1524043173028741, 0.0760894927419266,
         0.0077604214040978, 0.0024647284755382, 0.0780348340049386, 5.5059777270269628)
    Cc = (0, 0.0367762454319673, 0.1249685262725025, 0.2446177702277698, 0.2476149531070420, 0.2969311120382472,
          0.3978149645802642, 0.5270854589440328, 0.6981269994175695, 0.8190890835352128, 0.8527059887098624,
          0.8604711817462826, 0.8627060376969976, 0.8734213127600976)
    return tuple(q(x) for x in A), tuple(q(x) for x in B), tuple(q(x) for x in Cc)


def lsrk_euler(dtype=np.float64):
    q = _rt(dtype)
    return (q(0),), (q(1),), (q(0),)
