"""imc_refpy.py — a SECOND, independent restatement of the reference's transport step, in plain Python.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/test_refpy.py and by nothing else.

Why it exists.  The C++ oracle (imc_oracle.hpp) is itself a reading of the Julia sources, and Julia cannot run here
(SURVEY.md §8c), so nothing checks that reading.  This file reads the same Julia functions again, separately, in the
language whose scalar semantics are closest to Julia's: every value is a numpy scalar (np.float16 / np.float32 /
np.float64), so the reference's type promotions happen by themselves — `Float32 * Float64 -> Float64`, `Int * Float32 ->
Float32`, Float16 arithmetic computed in Float32 and rounded once — instead of being re-derived by hand as in the C++
template code.  Float64 literals of the Julia source are written `F64(...)` here because numpy treats bare Python floats
as "weak" (NEP 50), which Julia does not.  tests/test_refpy.py runs the C++ oracle and this file on the same decks,
particle batches and pre-drawn random numbers (replay tapes) and requires bit-identical particles, events and fields.
Pure-Python loops: small cases only (hundreds of particles).

Not independent: the elementary functions.  Julia's exp / expm1 / log / sin / cos / atan / ^ cannot be reproduced bit
for bit anyway, so both restatements call the same deterministic implementations (imc_math.h through the oracle's
`imc_oracle_math_eval` test hook); what is cross-checked is everything around them — control flow, event logic,
operation order, promotions, quirks Q1-Q32 of SURVEY.md §9.

Reference functions restated (file:line of /root/reference/src):
  sorter            imc_utilities.jl:23-54        update   imc_update.jl:12-70      clean  imc_clean.jl:6-19
  sourcing          imc_sourcing.jl:12-370        tally    imc_tally.jl:11-149      energychecker imc_energycheck.jl:10-37
  sample_planck     imc_sourcing.jl:372-399
  MC                imc_transport.jl:13-210       MC_RW    imc_transport.jl:212-479 MC2D   imc_transport.jl:483-732
  P_r / bisection / randomwalk_table              imc_transport.jl:734-797
"""
import ctypes as C
import math

import numpy as np

F64 = np.float64
_PREC_ID = {np.dtype(np.float16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}
_DP = C.POINTER(C.c_double)
np.seterr(all="ignore")


# ------------------------------------------------------------------------------------------------- Julia scalar helpers
def jtype(*xs):
    """Julia's promote_type of the arguments (Python ints are Julia Ints: they never widen a float)."""
    fl = [np.dtype(type(x)) for x in xs if isinstance(x, np.floating)]
    return np.result_type(*fl).type if fl else None


def jmin(*xs):
    """Base.min after promotion: NaN wins, -0.0 < +0.0."""
    T = jtype(*xs)
    r = T(xs[0])
    for x in xs[1:]:
        x = T(x)
        if np.isnan(r) or np.isnan(x):
            r = T(np.nan)
        elif x < r or (x == r and np.signbit(x)):
            r = x
    return r


def jmax(a, b):
    T = jtype(a, b)
    a, b = T(a), T(b)
    if np.isnan(a) or np.isnan(b):
        return T(np.nan)
    return b if (b > a or (a == b and np.signbit(a))) else a


def jl_sum(a):
    """Base.sum over a vector: mapreduce_impl, pairwise above 1024 elements (julia/base/reduce.jl)."""
    a = list(a)

    def rec(first, last):
        if first == last:
            return a[first]
        if last - first < 1024:
            v = a[first] + a[first + 1]
            for i in range(first + 2, last + 1):
                v = v + a[i]
            return v
        mid = first + ((last - first) >> 1)
        return rec(first, mid) + rec(mid + 1, last)
    return rec(0, len(a) - 1)


class JuliaMath:
    """exp, expm1, log, cos, sin, atan(y, x), x^y evaluated by imc_math.h in the precision of the (promoted) argument."""

    def __init__(self, oracle_dll):
        f = oracle_dll.imc_oracle_math_eval
        f.restype = C.c_int
        f.argtypes = [C.c_int32, C.c_int32, _DP, _DP, _DP, C.c_int64]
        self._f = f

    def _eval(self, fn, T, x, y=None):
        xa = np.array([float(x)]); out = np.empty(1)
        ya = None if y is None else np.array([float(y)])
        rc = self._f(fn, _PREC_ID[np.dtype(T)], xa.ctypes.data_as(_DP), None if ya is None else ya.ctypes.data_as(_DP), out.ctypes.data_as(_DP), 1)
        assert rc == 0
        return T(out[0])

    def exp(self, x): return self._eval(0, type(x), x)
    def expm1(self, x): return self._eval(1, type(x), x)
    def log(self, x): return self._eval(2, type(x), x)
    def sin(self, x): return self._eval(3, type(x), x)
    def cos(self, x): return self._eval(4, type(x), x)

    def atan(self, y, x):
        T = jtype(y, x)
        return self._eval(5, T, T(y), T(x))

    def pow(self, x, y):
        T = jtype(x, y)
        return self._eval(6, T, T(x), T(y))


class ReferenceThrows(Exception):
    """Base of the exceptions the Julia reference itself would raise at this point."""


class BoundsError(ReferenceThrows, IndexError):
    """The reference indexes past the end of an array here and Julia throws (e.g. mesh.dx[j] with j > Nx, Q7)."""


class InexactError(ReferenceThrows, ValueError):
    """Utilities.tointeger / Int(x) on a value that is not an integer (NaN, Inf, a fraction): Julia throws."""


class Tape:
    """Pre-drawn random numbers of one particle: rand(T) values and Float64 randexp() values, consumed in call order."""

    def __init__(self, uniforms, exponentials=()):
        self.u, self.e, self.iu, self.ie = list(uniforms), list(exponentials), 0, 0

    def rand(self, T):
        v = T(self.u[self.iu]); self.iu += 1
        return v

    def randexp64(self):
        v = F64(self.e[self.ie]); self.ie += 1
        return v

    def randexp(self, T):
        return T(self.randexp64())


# ------------------------------------------------------------------------------------------------- imc_utilities.jl
def sorter(vals, scales, T):
    """Utilities.sorter (imc_utilities.jl:23-54): pair the smallest with the largest factor, round every pair product to T."""
    for j, sc in enumerate(scales, start=1):
        E = jtype(*vals, sc)                          # eltype of vcat(vals, scales[j]); Int scales do not widen it
        s = sorted(E(v) for v in list(vals) + [sc])
        n = len(s)
        product = T(1.0)
        for i in range(n // 2):
            product = product * T(s[i] * s[n - 1 - i])
        if n % 2 == 1:
            product = product * T(s[n // 2])
        if not np.isinf(product) and not np.isnan(product):
            return product, sc, j
    return T(0.0), T(0.0), 0


# ------------------------------------------------------------------------------------------------- the state the stages share
class State:
    """mesh + simvars + constants of the reference, as Python lists of numpy scalars (1-based cell c lives at index c-1;
    2-D fields are dicts keyed by (xindex, yindex), 1-based)."""

    def __init__(self, inputs, mesh, simvars, consts, math_hook):
        self.T = T = inputs["PRECISION"]
        self.m = math_hook
        self.geometry = mesh.geometry
        self.linearized = str(inputs["LINEARIZED"]).upper() == "TRUE"
        self.name = str(inputs["NAME"]).upper()
        self.pairwise = str(simvars.pairwise).upper() == "TRUE"
        self.BC = tuple(simvars.BC)
        self.c, self.a, self.alpha = T(consts.phys_c), T(consts.phys_a), T(consts.alpha)
        self.ds = T(mesh.distancescale)
        self.scales = [T(s) for s in np.atleast_1d(mesh.energyscales)]
        self.n_input, self.n_max, self.cellmin = int(simvars.n_input), int(simvars.n_max), T(simvars.cellmin)
        self.dx = [T(v) for v in mesh.dx]
        if self.geometry == "1D":
            self.N = int(mesh.Ncells)
            self.cells = list(range(1, self.N + 1))                       # eachindex(mesh.temp)
            get = lambda arr: {c: T(arr[c - 1]) for c in self.cells}
            self.dy = None
            self.tsurf = [T(mesh.temp_surf[0]), T(mesh.temp_surf[1])]
        else:
            self.Nx, self.Ny = int(mesh.Ncells[0]), int(mesh.Ncells[1])
            self.cells = [(i, j) for j in range(1, self.Ny + 1) for i in range(1, self.Nx + 1)]   # column-major order
            get = lambda arr: {(i, j): T(arr[i - 1, j - 1]) for (i, j) in self.cells}
            self.dy = [T(v) for v in mesh.dy]
            self.tsurf = [[T(v) for v in side] for side in mesh.temp_surf]
        self.temp = get(mesh.temp)
        self.sa_const, self.sa_pow = get(mesh.sigma_a[..., 1]), get(mesh.sigma_a[..., 2])
        self.ss_const, self.ss_pow = get(mesh.sigma_s[..., 1]), get(mesh.sigma_s[..., 2])
        self.sigma_a, self.sigma_s = get(mesh.sigma_a[..., 0]), get(mesh.sigma_s[..., 0])
        self.sigma_static = get(mesh.sigma[..., 0])
        self.bee, self.radsource = get(mesh.bee), get(mesh.radsource)
        zero = lambda: {c: T(0.0) for c in self.cells}
        self.fleck, self.beta = zero(), zero()
        self.matenergydens, self.radenergydens = zero(), zero()
        self.energydep = {}      # (cell, scale index) -> T
        self.emittedenergy = {}
        self.totalenergy = T(0.0); self.totalenergydep = T(0.0); self.lostenergy = T(0.0); self.radenergyold = T(0.0)
        self.iterations = 0
        self.dt, self.t = T(simvars.dt), T(simvars.t)
        self.particles = []      # list of slot lists (numpy scalars; integer slots are stored as floats like in Julia)
        self.rw = None

    # helpers for the two geometries
    def w(self, c):              # dx[c] (1-D) — only used by 1-D code
        return self.dx[c - 1]

    def field(self, d):
        """dict -> flat float64 array in Julia's linear (column-major) order."""
        return np.array([float(d[c]) for c in self.cells])

    def field_scaled(self, d):
        return np.array([float(d.get((c, k), self.T(0.0))) for k in range(1, len(self.scales) + 1) for c in self.cells])

    def slots(self):
        return np.array([[float(v) for v in p] for p in self.particles]).reshape(len(self.particles), 9 if self.geometry == "1D" else 10)


# ------------------------------------------------------------------------------------------------- imc_update.jl
def update(S):
    """Update.update (imc_update.jl:12-70)."""
    T, m = S.T, S.m
    for c in S.cells:
        t = S.temp[c]
        if S.linearized:
            S.bee[c] = T((4 * S.a) * (t * t * t))           # .= into the T array              (:23)
            S.beta[c] = T(1.0)                               # ones(precision)                  (:24)
        else:
            S.beta[c] = ((4 * S.a) * (t * t * t)) / S.bee[c]  # a new array: Float64 when temp is (:26)
    for c in S.cells:
        t = S.temp[c]
        S.sigma_a[c] = T(S.sa_const[c] * m.pow(t, S.sa_pow[c]))                                  # (:30, :52)
        if S.geometry == "1D" and S.name == "MARSHAK WAVE":
            S.sigma_a[c] = T(((S.sa_const[c] / t) / t) / t)                                      # (:32-34)
        S.sigma_s[c] = T(S.ss_const[c] * m.pow(t, S.ss_pow[c]))                                  # (:35, :53)
    for c in S.cells:
        prod = sorter([S.ds, S.alpha, S.beta[c], S.c, S.dt, S.sigma_a[c]], [1], T)[0]
        S.fleck[c] = T(F64(1.0) / (F64(1.0) + prod))                                             # (:39, :58)


# ------------------------------------------------------------------------------------------------- imc_sourcing.jl
def _count(S, e, esc, n_source):
    """tointeger(max(round((e/escale)*n_source/totalenergy), cellmin)) stored into a zeros(T) array."""
    v = jmax(np.round(((e / esc) * n_source) / S.totalenergy), S.cellmin)
    if not np.isfinite(v) or float(v) != int(v):
        raise InexactError(f"tointeger({v!r})")
    return S.T(int(v))


def sourcing(S, tapes):
    """Sourcing.sourcing (imc_sourcing.jl:12-370).  `tapes` yields one Tape per new particle, in emission order."""
    T, ds, dt, c_, a_ = S.T, S.ds, S.dt, S.c, S.a
    tapes = iter(tapes)
    e_body, esc_body, e_rad, esc_rad = {}, {}, {}, {}
    S.emittedenergy = {}
    quarter = F64(0.25)
    if S.geometry == "1D":
        tl, tr = S.tsurf
        e_left, esc_left, _ = sorter([a_, c_, tl, tl, tl, tl, dt, quarter], S.scales, T)                       # (:60)
        e_right, esc_right, _ = sorter([a_, c_, tr, tr, tr, tr, dt, quarter], S.scales, T)                     # (:61)
        e_surface = (e_left / esc_left) + (e_right / esc_right)                                                # (:62)
        emitted_scale = {}
        for c in S.cells:
            t, f, sa, dx = S.temp[c], S.fleck[c], S.sigma_a[c], S.w(c)
            e_body[c], esc_body[c], _ = sorter([f, sa, a_, c_, t, t, t, t, dx, dt, ds], S.scales, T)            # (:67)
            e_body[c], esc_body[c] = T(e_body[c]), T(esc_body[c])                                               # stored into zeros(T) / ones(T)
            e_rad[c], esc_rad[c], _ = sorter([S.radsource[c], dx, dt], S.scales, T)                             # (:68)
            em, emitted_scale[c], k = sorter([f, sa, a_, c_, t, t, t, t, dt, ds], S.scales, T)                  # (:69)
            if k == 0:
                raise BoundsError("mesh.emittedenergy[i, 0]: no scale gives a finite product (imc_sourcing.jl:70)")
            S.emittedenergy[(c, k)] = T(em)
    else:
        eb, et, el, er = {}, {}, {}, {}
        sb, st, sl, sr = {}, {}, {}, {}
        for i in range(1, S.Nx + 1):
            tb, tt, dx = S.tsurf[0][i - 1], S.tsurf[1][i - 1], S.dx[i - 1]
            eb[i], sb[i], _ = sorter([a_, c_, tb, tb, tb, tb, dx, dt, quarter], S.scales, T)                    # (:87)
            et[i], st[i], _ = sorter([a_, c_, tt, tt, tt, tt, dx, dt, quarter], S.scales, T)                    # (:88)
        for j in range(1, S.Ny + 1):
            tl, tr, dy = S.tsurf[2][j - 1], S.tsurf[3][j - 1], S.dy[j - 1]
            el[j], sl[j], _ = sorter([a_, c_, tl, tl, tl, tl, dy, dt, quarter], S.scales, T)                    # (:91)
            er[j], sr[j], _ = sorter([a_, c_, tr, tr, tr, tr, dy, dt, quarter], S.scales, T)                    # (:92)
        side_sum = lambda e, s, n: jl_sum([T(e[k]) / T(s[k]) for k in range(1, n + 1)])
        e_surface = ((side_sum(eb, sb, S.Nx) + side_sum(et, st, S.Nx)) + side_sum(el, sl, S.Ny)) + side_sum(er, sr, S.Ny)   # (:95)
        for c in S.cells:
            i, j = c
            t, f, sa, dx, dy = S.temp[c], S.fleck[c], S.sigma_a[c], S.dx[i - 1], S.dy[j - 1]
            e_body[c], esc_body[c], _ = sorter([f, sa, a_, c_, t, t, t, t, dx, dy, dt, ds], S.scales, T)        # (:99)
            e_body[c], esc_body[c] = T(e_body[c]), T(esc_body[c])
            e_rad[c], esc_rad[c], _ = sorter([S.radsource[c], dx, dy, dt], S.scales, T)                         # (:100)
            em, _, k = sorter([f, sa, a_, c_, t, t, t, t, dt, ds], S.scales, T)                                 # (:101)
            if k == 0:
                raise BoundsError("mesh.emittedenergy[x, y, 0]: no scale gives a finite product (imc_sourcing.jl:102)")
            S.emittedenergy[(c, k)] = T(em)
    S.totalenergy = (jl_sum([e_body[c] / esc_body[c] for c in S.cells]) + jl_sum([e_rad[c] / esc_rad[c] for c in S.cells])) + e_surface   # (:121)

    n_source = S.n_input                                                                                        # (:131)
    n_census = len(S.particles)
    if S.n_input + n_census > S.n_max:
        cap = S.n_max - n_census - (1 if S.geometry == "1D" else 2) - 1                                         # length(Ncells)  (Q9)
        n_source = jmax(S.cellmin, T(cap))                                                                      # max(::T, ::Int) promotes to T
    n_body = {c: _count(S, e_body[c], esc_body[c], n_source) for c in S.cells}                                  # (:138-140)
    n_rad = {c: (_count(S, e_rad[c], esc_rad[c], n_source) if e_rad[c] > 0 else T(0.0)) for c in S.cells}       # (:142-146)

    def irange(n):                 # 1:n with a float upper bound
        return range(int(math.floor(float(n))))

    new = S.particles
    if S.geometry == "1D":
        N = S.N
        n_left = n_right = 0
        def surf_count(e, esc):
            v = np.round(T(((e / esc) * n_source) / S.totalenergy))                                             # round(precision, x)
            if not np.isfinite(v):
                raise InexactError(f"tointeger({v!r})")
            return int(v)
        if e_left > 0:
            n_left = surf_count(e_left, esc_left)                                                               # (:152)
        if e_right > 0:
            n_right = surf_count(e_right, esc_right)                                                            # (:156)
        for _ in range(n_left):                                                                                 # (:159-174)
            tp = next(tapes)
            xpos = T((F64(0.01) * S.w(1)) * ds)
            nrg = T(e_left / n_left)
            mu = T(np.sqrt(tp.rand(T)))
            while mu == 0.0:
                mu = T(np.sqrt(tp.rand(T)))
            spawn = dt * tp.rand(T)
            new.append([T(1), spawn, T(1), xpos, mu, T(1.0), nrg, nrg, T(esc_left)])
        for _ in range(n_right):                                                                                # (:175-189)
            tp = next(tapes)
            xpos = T((F64(0.99) * S.w(N)) * ds)
            nrg = T(e_right / n_right)
            mu = T(-np.sqrt(tp.rand(T)))
            while mu == 0.0:
                mu = T(-np.sqrt(tp.rand(T)))
            spawn = dt * tp.rand(T)
            new.append([T(N), spawn, T(N), xpos, mu, T(1.0), nrg, nrg, T(esc_right)])
        for e, esc, cnt in ((e_body, esc_body, n_body), (e_rad, esc_rad, n_rad)):                               # (:192-212), (:214-233)
            for c in S.cells:
                if cnt[c] <= 0:
                    continue
                nrg = T(e[c]) / T(cnt[c])
                for _ in irange(cnt[c]):
                    tp = next(tapes)
                    xpos = (S.w(c) * tp.rand(T)) * ds
                    mu = T(1 - 2 * tp.rand(T))
                    while mu == 0.0:
                        mu = T(1 - 2 * tp.rand(T))
                    spawn = dt * tp.rand(T)
                    new.append([T(c), spawn, T(c), xpos, mu, T(1.0), nrg, nrg, T(esc[c])])
    else:
        Nx, Ny = S.Nx, S.Ny
        pi64, piT = F64(math.pi), T(math.pi)                    # Irrational * T -> T(pi) * x;  -pi, 2*pi, pi*Float64 -> Float64
        cnt_side = lambda e, s, n: {k: (_count(S, T(e[k]), T(s[k]), n_source) if e[k] > 0 else T(0.0)) for k in range(1, n + 1)}
        nb, nt, nl, nr = cnt_side(eb, sb, Nx), cnt_side(et, st, Nx), cnt_side(el, sl, Ny), cnt_side(er, sr, Ny)   # (:240-263)
        for i in range(1, Nx + 1):                                                                              # bottom (:265-278)
            for _ in irange(nb[i]):
                tp = next(tapes)
                spawn = dt * tp.rand(T)
                xpos = (S.dx[i - 1] * tp.rand(T)) * ds
                ypos = T((F64(0.001) * S.dy[0]) * ds)
                mu = T(piT * tp.rand(T))
                nrg = T(eb[i]) / nb[i]
                new.append([spawn, T(i), T(1), xpos, ypos, mu, T(1.0), nrg, nrg, T(sb[i])])
        for i in range(1, Nx + 1):                                                                              # top (:280-293)
            for _ in irange(nt[i]):
                tp = next(tapes)
                spawn = dt * tp.rand(T)
                xpos = (S.dx[i - 1] * tp.rand(T)) * ds
                ypos = T((F64(0.999) * S.dy[Ny - 1]) * ds)
                mu = T((-pi64) * tp.rand(T))
                nrg = T(et[i]) / nt[i]
                new.append([spawn, T(i), T(Ny), xpos, ypos, mu, T(1.0), nrg, nrg, T(st[i])])
        for j in range(1, Ny + 1):                                                                              # left (:295-308)
            for _ in irange(nl[j]):
                tp = next(tapes)
                spawn = dt * tp.rand(T)
                xpos = T((F64(0.001) * S.dx[0]) * ds)
                if j > Nx:
                    raise BoundsError("mesh.dx[j] with j > Nx (imc_sourcing.jl:301)")
                ypos = (S.dx[j - 1] * tp.rand(T)) * ds                                                          # dx indexed by j (Q7)
                mu = T(pi64 * (F64(0.5) - tp.rand(T)))
                nrg = T(el[j]) / nl[j]
                new.append([spawn, T(1), T(j), xpos, ypos, mu, T(1.0), nrg, nrg, T(sl[j])])
        for j in range(1, Ny + 1):                                                                              # right (:310-323)
            for _ in irange(nr[j]):
                tp = next(tapes)
                spawn = dt * tp.rand(T)
                xpos = T((F64(0.999) * S.dx[Nx - 1]) * ds)
                if j > Nx:
                    raise BoundsError("mesh.dx[j] with j > Nx (imc_sourcing.jl:316)")
                ypos = (S.dx[j - 1] * tp.rand(T)) * ds
                mu = T(pi64 * (F64(0.5) + tp.rand(T)))
                nrg = T(er[j]) / nr[j]
                new.append([spawn, T(Nx), T(j), xpos, ypos, mu, T(1.0), nrg, nrg, T(sr[j])])
        for e, esc, cnt in ((e_body, esc_body, n_body), (e_rad, esc_rad, n_rad)):                               # (:325-345), (:347-366)
            for c in S.cells:
                i, j = c
                if cnt[c] <= 0:
                    continue
                nrg = T(e[c]) / T(cnt[c])
                for _ in irange(n_body[c]):                                                                     # radsource loops 1:n_body too (Q6)
                    tp = next(tapes)
                    xpos = (S.dx[i - 1] * tp.rand(T)) * ds
                    ypos = (S.dy[j - 1] * tp.rand(T)) * ds
                    mu = T((2 * pi64) * tp.rand(T))
                    spawn = dt * tp.rand(T)
                    new.append([spawn, T(i), T(j), xpos, ypos, mu, T(1.0), nrg, nrg, T(esc[c])])


def sample_planck(T, tp, m, max_terms=100000):
    """Sourcing.sample_planck (imc_sourcing.jl:372-399).  `while true` in the reference; a first draw above the largest
    value 90 nsum / pi^4 reaches in T never leaves it — NaN after max_terms by the convention of include/imc.h."""
    n = T(1.0)
    rn1 = tp.rand(T)
    nsum = T(1.0)
    pi = F64(math.pi)
    pi4 = (pi * pi) * (pi * pi)                       # pi^4: Float64 power by squaring
    for _ in range(max_terms):
        if rn1 <= F64(90.0) * nsum / pi4:                                                                       # (:388)
            rn1 = tp.rand(T); rn2 = tp.rand(T); rn3 = tp.rand(T); rn4 = tp.rand(T)
            return T(F64(-1.0) * m.log(rn1 * rn2 * rn3 * rn4) / n)                                              # (:393)
        n = n + T(1.0)                                                                                          # (:396)
        n4 = T(np.float32(n) ** 4) if T is np.float16 else (n * n) * (n * n)                                    # Float16: Float32(n)^4, rounded once
        nsum = nsum + T(F64(1.0) / n4)                                                                          # (:397)
    return T(np.nan)


# ------------------------------------------------------------------------------------------------- imc_transport.jl
class _Deposits:
    """energydep / lostenergy accumulation of MC, MC_RW and MC2D: `+=` in T, or per-cell vectors summed with Base.sum."""

    def __init__(self, S):
        self.S = S
        S.energydep = {}
        self.vec, self.lost = {}, {}

    def dep(self, cell, k, v):
        S = self.S
        if S.pairwise:
            self.vec.setdefault((cell, k), []).append(S.T(v))        # push! into a Vector{precision}
        else:
            S.energydep[(cell, k)] = S.T(S.energydep.get((cell, k), S.T(0.0)) + v)

    def lose(self, k, energy, energyscale):
        S = self.S
        if S.pairwise:
            self.lost.setdefault(k, []).append(S.T(energy))
        else:
            S.lostenergy = S.lostenergy + energy / energyscale

    def finish(self):
        S = self.S
        if S.pairwise:
            for k in range(1, len(S.scales) + 1):
                lv = self.lost.get(k, [])
                S.lostenergy = S.lostenergy + (jl_sum(lv) if lv else S.T(0.0)) / S.scales[k - 1]
                for c in S.cells:
                    v = self.vec.get((c, k), [])
                    S.energydep[(c, k)] = jl_sum(v) if v else S.T(0.0)


def _scale_index(S, energyscale):
    for k, s in enumerate(S.scales, start=1):       # findfirst(isequal(energyscale), mesh.energyscales)
        if s == energyscale:
            return k
    raise KeyError(energyscale)


def MC(S, tapes, outcomes=None):
    """Transport.MC (imc_transport.jl:13-210).  tapes[i] feeds particle i.  outcomes: optional list receiving (event, nseg)."""
    T, m, ds, c_ = S.T, S.m, S.ds, S.c
    D = _Deposits(S)
    for idx, p in enumerate(S.particles):
        tp = tapes[idx]
        origin, time, cell, x, mu, freq, energy, e0, escale = p[0], p[1], int(p[2]), p[3], p[4], p[5], p[6], p[7], p[8]
        k = _scale_index(S, escale)
        emin = T(F64(0.01) * e0)                                                                                # (:61)
        nseg, event = 0, None
        while True:
            S.iterations += 1; nseg += 1                                                                         # (:73)
            dx, sa, ss, f = S.w(cell), S.sigma_a[cell], S.sigma_s[cell], S.fleck[cell]
            if mu > 0.0:
                d_b = (dx * ds - x) / mu                                                                         # (:79)
            else:
                d_b = abs(x / mu)                                                                                # (:82)
            d_col = tp.randexp(T) / (sa * (1 - f) + ss)                                                          # (:87)
            d_cen = (c_ * (S.dt - time)) * ds                                                                    # (:89)
            d = jmin(d_b, d_col, d_cen)                                                                          # (:92)
            e_new = energy * m.exp(((-sa) * f) * d)                                                              # (:95)
            if e_new <= emin:                                                                                    # (:97-106)
                D.dep(cell, k, energy / dx)
                p[7] = T(-1.0)
                event = 1
                break
            D.dep(cell, k, (-(energy / dx)) * m.expm1(((-f) * sa) * d))                                          # (:110 / :120)
            x = x + mu * d                                                                                       # (:124)
            time = time + (d / ds) / c_                                                                          # (:125)
            energy = e_new
            if d == d_b:                                                                                         # (:130-171)
                if mu > 0:
                    if cell == S.N:
                        if S.BC[1] == "REFLECT":
                            mu = -mu
                        elif S.BC[1] == "VACUUM":
                            D.lose(k, energy, escale)
                            p[7] = T(-1.0)
                            event = 2
                            break
                    cell += 1
                    x = 0                                                                                        # an Int in Julia; arithmetic treats it as 0
                if mu < 0:
                    if cell == 1:
                        if S.BC[0] == "REFLECT":
                            mu = -mu
                        elif S.BC[0] == "VACUUM":
                            D.lose(k, energy, escale)
                            p[7] = T(-1.0)
                            event = 2
                            break
                    else:
                        cell -= 1
                        x = S.w(cell) * ds
            if d == d_col:                                                                                       # (:174-183)
                mu = T(0.0)
                while mu == 0.0:
                    mu = T(1 - 2 * tp.rand(T))
            if d == d_cen:                                                                                       # (:185-193)
                p[:] = [T(v) for v in (origin, T(0.0), cell, x, mu, freq, energy, e0, escale)]
                event = 0
                break
        if outcomes is not None:
            outcomes.append((event, nseg))
    D.finish()                                                                                                   # (:197-205)


def P_r(S, a):
    """Transport.P_r (imc_transport.jl:734-754): T(1.0) for a == 0, otherwise a Float64 partial sum of 100 terms."""
    if a != 0:
        Pr = S.T(0.0)
        for n in range(1, 101):
            pin = F64(math.pi) * n
            Pr = Pr + ((-1) ** (n - 1)) * S.m.exp((-a) * (pin * pin)) * 2
        return Pr
    return S.T(1.0)


def bisection(array, value):
    """Transport.bisection (imc_transport.jl:756-784), 1-based result."""
    n = len(array)
    if value < array[0]:
        return 1
    elif value > array[n - 1]:
        return n
    jl_, ju = 1, n
    while ju - jl_ > 1:
        jm = (ju + jl_) >> 1
        if value >= array[jm - 1]:
            jl_ = jm
        else:
            ju = jm
    if value == array[0]:
        return 1
    elif value == array[n - 1]:
        return n
    return jl_


def randomwalk_table(S, lo=0, hi=10, n=1000):
    """main's table set-up (MixedPrecisionIMC.jl:129-133) + Transport.randomwalk_table (imc_transport.jl:786-797)."""
    T = S.T
    aVals = [T((1.0 - i / (n - 1)) * lo + (i / (n - 1)) * hi) for i in range(n)]      # precision.(LinRange(0, 10, 1000))
    prVals, ptVals = [], []
    for a in aVals:
        prVals.append(T(P_r(S, a)))             # stored into zeros(precision)
        ptVals.append(T(1 - prVals[-1]))
    S.rw = (aVals, prVals, ptVals)


def MC_RW(S, tapes, outcomes=None):
    """Transport.MC_RW (imc_transport.jl:212-479).  Ignores distancescale and draws randexp() in Float64 (Q3)."""
    T, m, c_ = S.T, S.m, S.c
    aVals, prVals, ptVals = S.rw
    D = _Deposits(S)
    for idx, p in enumerate(S.particles):
        tp = tapes[idx]
        origin, time, cell, x, mu, freq, energy, e0, escale = p[0], p[1], int(p[2]), p[3], p[4], p[5], p[6], p[7], p[8]
        k = _scale_index(S, escale)
        emin = T(F64(0.01) * e0)
        nseg, event = 0, None
        while True:
            S.iterations += 1; nseg += 1                                                                         # (:265)
            dx, sa, ss, f = S.w(cell), S.sigma_a[cell], S.sigma_s[cell], S.fleck[cell]
            if mu > 0.0:
                d_b = (dx - x) / mu                                                                              # (:271)
            else:
                d_b = abs(x / mu)
            d_col = abs(tp.randexp64()) / (sa * (1 - f) + ss)                                                    # (:279)
            d_cen = c_ * (S.dt - time)                                                                           # (:281)
            d = jmin(d_b, d_col, d_cen)
            R0 = jmin(abs(dx - x), abs(x))                                                                       # (:287)
            if R0 > 1 / S.sigma_static[cell] and d_col < R0:                                                     # (:289)  (Q4)
                u = tp.rand(T)
                Dc = T(c_ / ((3 * sa) * (1 - f)))                                                                # (:292)
                a = (Dc * S.dt) / (R0 * R0)
                Pr = P_r(S, a)
                Pt = 1 - Pr
                if u < Pt:                                                                                       # (:298)
                    ai = bisection(ptVals, u)
                    t_p = T((aVals[ai - 1] * (R0 * R0)) / Dc)
                    arg = (((t_p * c_) * (1 - f)) * sa) / m.log(1 - f)
                    e_new = energy * m.exp(arg)
                    if e_new <= e0:                                                                              # always true (Q1)
                        e_new = F64(0.0)
                    D.dep(cell, k, (-(energy / dx)) * m.expm1(arg))                                              # (:317 / :319)
                    if e_new == 0.0:
                        p[7] = T(-1.0); event = 3
                        break
                    raise AssertionError("unreachable in the reference (Q1)")
                else:
                    u_prime = tp.rand(T)                                                                         # (:338)
                    Pr0 = Pr * u_prime
                    ai = bisection(prVals, Pr0 * u_prime)                                                        # decreasing table (Q5)
                    _R1 = np.sqrt((Dc * S.dt) / aVals[ai - 1])
                    arg = (((c_ * (1 - f)) * sa) * S.dt) / m.log(1 - f)
                    e_new = energy * m.exp(arg)
                    if e_new <= e0:
                        e_new = F64(0.0)
                    D.dep(cell, k, (-(energy / dx)) * m.expm1(arg))                                              # (:353 / :355)
                    if e_new == 0.0:
                        p[7] = T(-1.0); event = 3
                        break
                    raise AssertionError("unreachable in the reference (Q1)")
            e_new = energy * m.exp(((-sa) * f) * d)                                                              # (:374)
            if e_new <= emin:
                e_new = T(0.0)
            D.dep(cell, k, energy - e_new)                                                                       # not divided by dx (Q2)
            if e_new == 0.0:
                p[7] = T(-1.0); event = 1
                break
            x = x + mu * d                                                                                       # (:397-399)
            time = time + d / c_
            energy = e_new
            if d == d_b:
                if mu > 0:
                    if cell == S.N:
                        if S.BC[1] == "REFLECT":
                            mu = -mu
                        elif S.BC[1] == "VACUUM":
                            D.lose(k, energy, escale)
                            p[7] = T(-1.0); event = 2
                            break
                    cell += 1
                    x = T(0)
                if mu < 0:
                    if cell == 1:
                        if S.BC[0] == "REFLECT":
                            mu = -mu
                        elif S.BC[0] == "VACUUM":
                            D.lose(k, energy, escale)
                            p[7] = T(-1.0); event = 2
                            break
                    else:
                        cell -= 1
                        x = S.w(cell)
            if d == d_col:                                                                                       # (:447-453)
                mu = T(1 - 2 * tp.rand(T))
                while mu == 0.0:
                    mu = T(1 - 2 * tp.rand(T))
            if d == d_cen:                                                                                       # (:455-461): the slot vector is a Vector{T}
                p[:] = [T(v) for v in (origin, T(0.0), cell, x, mu, freq, energy, e0, escale)]
                event = 0
                break
        if outcomes is not None:
            outcomes.append((event, nseg))
    D.finish()


def MC2D(S, tapes, outcomes=None):
    """Transport.MC2D (imc_transport.jl:483-732)."""
    T, m, ds, c_ = S.T, S.m, S.ds, S.c
    D = _Deposits(S)
    for idx, p in enumerate(S.particles):
        tp = tapes[idx]
        time, xi, yi, x, y, mu, frq, energy, e0, escale = p[0], int(p[1]), int(p[2]), p[3], p[4], p[5], p[6], p[7], p[8], p[9]
        k = _scale_index(S, escale)
        emin = T(F64(0.01) * e0)                                                                                 # (:531)
        nseg, event = 0, None
        while True:
            nseg += 1; S.iterations += 1      # the reference does not count here (Q13); SURVEY §8d defines the count identically
            cell = (xi, yi)
            dxc, dyc, sa, ss, f = S.dx[xi - 1], S.dy[yi - 1], S.sigma_a[cell], S.sigma_s[cell], S.fleck[cell]
            vx, vy = T(m.cos(mu)), T(m.sin(mu))                                                                  # (:534)
            d_bx = abs((dxc * ds - x) / vx) if vx > 0 else abs(x / vx)                                           # (:536-540)
            d_by = abs((dyc * ds - y) / vy) if vy > 0 else abs(y / vy)                                           # (:542-546)
            if np.isnan(d_bx):                                                                                   # (:548-554)
                d_b = d_by
            elif np.isnan(d_by):
                d_b = d_bx
            else:
                d_b = jmin(d_bx, d_by)
            d_col = tp.randexp(T) / (sa * (1 - f) + ss)                                                          # (:561)
            d_cen = (c_ * (S.dt - time)) * ds                                                                    # (:568)
            d = jmin(d_b, d_col, d_cen)                                                                          # (:570)
            e_new = energy * m.exp(((-f) * sa) * d)                                                              # (:577)
            if e_new <= emin:                                                                                    # (:586-595)
                D.dep(cell, k, (energy / dxc) / dyc)
                p[7] = T(-1.0)
                event = 1
                break
            D.dep(cell, k, ((-(energy / dxc)) / dyc) * m.expm1(((-f) * sa) * d))                                 # (:598 / :601)
            x = x + d * vx                                                                                       # (:608-611)
            y = y + d * vy
            time = time + (d / ds) / c_
            energy = e_new
            if d == d_bx or d == d_by:                                                                           # (:621)
                if d_bx < d_by:
                    side, idx_, last, bc, flip = (m.cos(mu) > 0), xi, S.Nx, (S.BC[1], S.BC[0]), "x"
                else:
                    side, idx_, last, bc, flip = (m.sin(mu) > 0), yi, S.Ny, (S.BC[2], S.BC[3]), "y"
                at_wall = (idx_ == last) if side else (idx_ == 1)
                if at_wall:
                    which = bc[0] if side else bc[1]
                    if which == "REFLECT":                                                                       # xvec .- v .* (2(v'xvec)), atan(y, x)
                        rx = vx - 1 * (2 * vx) if flip == "x" else vx - 0 * (2 * vy)
                        ry = vy - 0 * (2 * vx) if flip == "x" else vy - 1 * (2 * vy)
                        mu = m.atan(ry, rx)
                    elif which == "VACUUM":
                        D.lose(k, energy, escale)
                        p[7] = T(-1.0)
                        event = 2
                        break
                    continue
                if flip == "x":
                    if side:
                        xi += 1; x = T(0.0)
                    else:
                        xi -= 1; x = S.dx[xi - 1] * ds
                else:
                    if side:
                        yi += 1; y = T(0.0)
                    else:
                        yi -= 1; y = S.dy[yi - 1] * ds
                continue
            if d == d_col:                                                                                       # (:707-709)
                mu = T((2 * F64(math.pi)) * tp.rand(T))
            if d == d_cen:                                                                                       # (:711-716)
                p[:] = [T(v) for v in (T(0), xi, yi, x, y, mu, frq, energy, e0, escale)]
                event = 0
                break
        if outcomes is not None:
            outcomes.append((event, nseg))
    D.finish()


# ------------------------------------------------------------------------------------------------- imc_clean.jl
def clean(S):
    """Clean.clean (imc_clean.jl:6-19): delete, from the back, every particle whose slot 8 is -1.0."""
    for i in range(len(S.particles) - 1, -1, -1):
        if S.particles[i][7] == -1.0:
            del S.particles[i]


# ------------------------------------------------------------------------------------------------- imc_tally.jl
def tally(S):
    """Tally.tally (imc_tally.jl:11-149).  Returns nrg_inc (what is pushed onto energyincrease_saved)."""
    T, m = S.T, S.m
    ns = len(S.scales)
    if S.t == 0.0:                                                                                               # (:29-33)  (Q11)
        for c in S.cells:
            t = S.temp[c]
            S.matenergydens[c] = T(sorter([S.fleck[c], S.sigma_a[c], S.a, S.c, t, t, t, t, S.dt, S.ds], [1], T)[0])
    S.totalenergydep = T(0.0)                                                                                    # (:43)
    nrg_inc = {c: T(0.0) for c in S.cells}
    for k in range(1, ns + 1):                                                                                   # (:46-57)
        sc = S.scales[k - 1]
        dep = {c: S.energydep.get((c, k), T(0.0)) for c in S.cells}
        em = {c: S.emittedenergy.get((c, k), T(0.0)) for c in S.cells}
        for c in S.cells:
            nrg_inc[c] = nrg_inc[c] + (dep[c] - em[c]) / sc
        if S.geometry == "1D":
            S.totalenergydep = S.totalenergydep + jl_sum([(dep[c] * S.w(c)) / sc for c in S.cells])
        else:
            S.totalenergydep = S.totalenergydep + jl_sum([(dep[c] * (S.dx[c[0] - 1] * S.dy[c[1] - 1])) / sc for c in S.cells])
    for c in S.cells:
        S.matenergydens[c] = S.matenergydens[c] + nrg_inc[c]                                                     # (:68)
    for c in S.cells:
        if S.linearized:
            S.temp[c] = m.pow(S.matenergydens[c], F64(1) / F64(4))                                               # Float64 from here on (:72, Q12)
        else:
            S.temp[c] = S.temp[c] + nrg_inc[c] / S.bee[c]                                                        # (:74)
    vec = {c: [] for c in S.cells}                                                                               # (:84-113)  (Q19)
    for p in S.particles:
        if S.geometry == "1D":
            c = int(p[2])
            vec[c].append(T(p[6] / (S.w(c) * p[8])))
        else:
            c = (int(p[1]), int(p[2]))
            vec[c].append(T(p[7] / ((S.dx[c[0] - 1] * S.dy[c[1] - 1]) * p[9])))
    for c in S.cells:
        S.radenergydens[c] = jl_sum(vec[c]) if vec[c] else T(0.0)
    return nrg_inc


# ------------------------------------------------------------------------------------------------- imc_energycheck.jl
def energychecker(S):
    """EnergyCheck.energychecker (imc_energycheck.jl:10-37).  Returns (radenergy, energy_error)."""
    if S.geometry == "1D":
        rad = jl_sum([S.radenergydens[c] * S.w(c) for c in S.cells])
    else:
        rad = jl_sum([(S.radenergydens[c] * S.dx[c[0] - 1]) * S.dy[c[1] - 1] for c in S.cells])
    err = (((S.totalenergy - S.totalenergydep) - (rad - S.radenergyold)) - S.lostenergy) / S.totalenergy       # (:34)
    S.radenergyold = rad
    S.lostenergy = S.T(0.0)
    return rad, err
