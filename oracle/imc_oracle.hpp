// imc_oracle.hpp — CPU restatement of the MixedPrecisionIMC.jl transport step.
//
// *** TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
// *** reference arm may build, load or call this.  The product (libimc_b200.so) never does.
//
// PARITY UNPINNED: Julia is not installed here and the reference's tests hold no
// golden vector for sourcing / MC / MC_RW / MC2D / tally, so this restatement is pinned only by
//   - the one reference KAT that touches the path (clean: test/runtests.jl:78-87),
//   - the deck-embedded Su-Olson benchmark curve (src/inputs/SuOlson.txt:71-72),
//   - the analytic infinite-medium equilibrium and energy conservation (imc_energycheck.jl:34),
// and by line-by-line reading of the Julia source, cited at each function.  That reading is cross-checked
// by a second restatement written separately in plain Python (oracle/imc_refpy.py; tests/test_refpy.py
// requires the two to agree bit for bit, stage by stage, on replay tapes).
// Julia's RNG stream and libm are not reproducible here: random numbers come from a tape
// (replay) or from Philox (csrc/imc_rng.h); elementary functions come from a Math policy:
// MathDet (csrc/imc_math.h, bit-identical to the GPU) or MathLibm (glibc, independent check).
//
// The code deliberately follows the reference's data structures (one small slot vector per
// particle, per-cell push! vectors for PAIRWISE) and statement order, including the quirks in
// SURVEY.md §9 (Qn tags below).  It is single-threaded like the reference.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <array>
#include <algorithm>

#include "imc.h"
#include "imc_num.h"
#include "imc_math.h"
#include "imc_rng.h"

namespace imc_oracle {
using namespace imc;

// ---- independent math policy: glibc (correctly rounded in practice), for cross-checks ----
struct MathLibm {
  template <class P> static Num<P> wrap(double r) { return Num<P>::from_d(r); }
  template <class P> static Num<P> exp(Num<P> x) { if constexpr (P::id == 2) return Num<P>(std::exp(x.v)); else return Num<P>(P::rnd(std::exp(x.v))); }
  template <class P> static Num<P> expm1(Num<P> x) { if constexpr (P::id == 2) return Num<P>(std::expm1(x.v)); else return Num<P>(P::rnd(std::expm1(x.v))); }
  template <class P> static void exp_expm1(Num<P> x, Num<P>* e, Num<P>* m) { *e = exp(x); *m = expm1(x); }
  template <class P> static Num<P> log(Num<P> x) { if constexpr (P::id == 2) return Num<P>(std::log(x.v)); else return Num<P>(P::rnd(std::log(x.v))); }
  template <class P> static Num<P> sqrt(Num<P> x) { if constexpr (P::id == 2) return Num<P>(std::sqrt(x.v)); else return Num<P>(P::rnd(std::sqrt(x.v))); }
  template <class P> static void sincos(Num<P> x, Num<P>* s, Num<P>* c) {
    if constexpr (P::id == 2) { *s = Num<P>(std::sin(x.v)); *c = Num<P>(std::cos(x.v)); }
    else { *s = Num<P>(P::rnd(std::sin(x.v))); *c = Num<P>(P::rnd(std::cos(x.v))); }
  }
  template <class P> static Num<P> atan2(Num<P> y, Num<P> x) { if constexpr (P::id == 2) return Num<P>(std::atan2(y.v, x.v)); else return Num<P>(P::rnd(std::atan2(y.v, x.v))); }
  template <class P> static Num<P> pow(Num<P> x, Num<P> y) { if constexpr (P::id == 2) return Num<P>(std::pow(x.v, y.v)); else return Num<P>(P::rnd((float)std::pow((double)x.v, (double)y.v))); }
  static double exp64(double x) { return std::exp(x); }
  static double expm164(double x) { return std::expm1(x); }
  static double log64(double x) { return std::log(x); }
  static double pow64(double x, double y) { return std::pow(x, y); }
  static double sqrt64(double x) { return std::sqrt(x); }
  static double cos64(double x) { return std::cos(x); }
};

// Julia Base.sum over a Vector{T}: mapreduce_impl pairwise with 1024-element sequential blocks
// (base/reduce.jl).  The block loop carries @simd in Julia, whose reassociation is code-generation
// dependent; the restatement uses strict left-to-right order inside a block (Q20).
template <class P>
Num<P> jl_sum_range(const Num<P>* a, size_t first, size_t last) {  // inclusive, first <= last
  if (first == last) return a[first];
  if (last - first < 1024) {
    Num<P> v = a[first] + a[first + 1];
    for (size_t i = first + 2; i <= last; ++i) v = v + a[i];
    return v;
  }
  size_t mid = first + ((last - first) >> 1);
  Num<P> v1 = jl_sum_range(a, first, mid);
  Num<P> v2 = jl_sum_range(a, mid + 1, last);
  return v1 + v2;
}
template <class P>
Num<P> jl_sum(const std::vector<Num<P>>& a) {
  if (a.empty()) return Num<P>();
  return jl_sum_range(a.data(), 0, a.size() - 1);
}

// Utilities.sorter (imc_utilities.jl:23-54).  `vals` are the Float64 images of the literal array's
// elements (exact for T values; genuinely Float64 where the reference leaks, Q12/Q31): the pair
// product is formed in the array's element type and converted to T, which for T-valued inputs
// equals one rounding of the exact product, i.e. from_d(a*b) in both cases.
template <class P>
struct SorterResult { Num<P> product; Num<P> scale; int index; };  // index is 1-based; 0 = failure
template <class P>
SorterResult<P> sorter(const double* vals, int n, const Num<P>* scales, int n_scales) {
  for (int j = 0; j < n_scales; ++j) {
    double sorted[24];
    int m = n + 1;
    for (int i = 0; i < n; ++i) sorted[i] = vals[i];
    sorted[n] = scales[j].d();
    for (int i = 1; i < m; ++i) {  // insertion sort, ascending
      double key = sorted[i];
      int k = i - 1;
      while (k >= 0 && sorted[k] > key) { sorted[k + 1] = sorted[k]; --k; }
      sorted[k + 1] = key;
    }
    Num<P> product = Num<P>::from_d(1.0);
    for (int i = 0; i < m / 2; ++i) product *= Num<P>::from_d(sorted[i] * sorted[m - 1 - i]);
    if (m & 1) product *= Num<P>::from_d(sorted[m / 2]);
    if (!is_inf(product) && !is_nan(product)) return {product, scales[j], j + 1};
  }
  return {Num<P>(), Num<P>(), 0};
}

// A value whose Julia type is either T or Float64 at run time (MC_RW, Q3).
template <class P>
struct Dyn {
  double v; bool wide;
  Dyn() : v(0), wide(false) {}
  Dyn(Num<P> x) : v(x.d()), wide(false) {}
  static Dyn w(double x) { Dyn r; r.v = x; r.wide = true; return r; }
  Num<P> narrow() const { return Num<P>::from_d(v); }  // exact when !wide
};
template <class P, class Op64, class OpT>
Dyn<P> dyn_op(Dyn<P> a, Dyn<P> b, Op64 f64, OpT ft) {
  if (a.wide || b.wide) return Dyn<P>::w(f64(a.v, b.v));
  return Dyn<P>(ft(a.narrow(), b.narrow()));
}
template <class P> Dyn<P> operator+(Dyn<P> a, Dyn<P> b) { return dyn_op(a, b, [](double x, double y) { return x + y; }, [](Num<P> x, Num<P> y) { return x + y; }); }
template <class P> Dyn<P> operator-(Dyn<P> a, Dyn<P> b) { return dyn_op(a, b, [](double x, double y) { return x - y; }, [](Num<P> x, Num<P> y) { return x - y; }); }
template <class P> Dyn<P> operator*(Dyn<P> a, Dyn<P> b) { return dyn_op(a, b, [](double x, double y) { return x * y; }, [](Num<P> x, Num<P> y) { return x * y; }); }
template <class P> Dyn<P> operator/(Dyn<P> a, Dyn<P> b) { return dyn_op(a, b, [](double x, double y) { return x / y; }, [](Num<P> x, Num<P> y) { return x / y; }); }
template <class P> Dyn<P> dabs(Dyn<P> a) { Dyn<P> r = a; r.v = std::fabs(a.v); return r; }
template <class P> Dyn<P> dmin(Dyn<P> a, Dyn<P> b) {  // promote, then Julia min
  Dyn<P> r; r.wide = a.wide || b.wide;
  r.v = jl_min(Num<F64>(a.v), Num<F64>(b.v)).v;
  return r;
}

struct EngineBase {
  std::string err;
  virtual ~EngineBase() {}
  virtual int set_mesh(const double*, const double*, const double*, const double*, const double*, const double*,
                       const double*, const double*, const double*, const double*, const double*, const double*,
                       const double*, const double*) = 0;
  virtual int rw_table(double, double, int, double*, double*, double*) = 0;
  virtual int update(double dt) = 0;
  virtual int source(double dt, int64_t n_input, double cellmin, int64_t step, int64_t n_census_global, imc_source_stats*) = 0;
  virtual int transport(double dt, int64_t step, imc_transport_stats*) = 0;
  virtual int clean(int64_t*) = 0;
  virtual int tally_local() = 0;
  virtual int tally_finish(double t, double dt, imc_tally_stats*) = 0;
  virtual int energycheck(imc_energy_stats*) = 0;
  virtual int reduce_buffer(void**, int64_t*, int32_t*) = 0;
  virtual int get_field(int, double*, int64_t) = 0;
  virtual int set_state(const double*, const double*, const double*) = 0;
  virtual int field_elsize(int) = 0;
  virtual int get_field_native(int, void*, int64_t) = 0;
  virtual int set_state_native(const void*, const void*, const void*) = 0;
  virtual int history_enable(int64_t) = 0;
  virtual int history_count(int64_t*, int64_t*) = 0;
  virtual int history_get(int, int64_t, int64_t, void*, int64_t) = 0;
  virtual int history_clear() = 0;
  virtual int64_t num_particles() = 0;
  virtual int get_particles(double*, uint64_t*, int64_t) = 0;
  virtual int set_particles(const double*, const uint64_t*, int64_t) = 0;
  virtual int set_transport_tape(const double*, int, const double*, int, int64_t) = 0;
  virtual int set_source_tape(const double*, int, int64_t) = 0;
  virtual int get_outcomes(int32_t*, int32_t*, int64_t) = 0;
  virtual int sample_planck(int64_t, int64_t, double*) = 0;
  virtual int checkpoint(int) = 0;
};

template <class P, class M>
struct Oracle : EngineBase {
  using N = Num<P>;
  using Slots = std::array<N, 10>;  // 9 used in 1-D, 10 in 2-D (reference layouts)
  imc_config cfg;
  int geom, nx, ny, nslots;
  size_t nc;
  int ns;
  std::vector<N> scales;
  N ds, phys_c, phys_a, alpha;
  // MeshStruct (imc_mesh.jl:10-40)
  std::vector<N> dx, dy;
  std::vector<double> temp; bool temp_wide = false;  // Q12: Float64 after the first LINEARIZED tally
  N tsurf1d[2];
  std::vector<N> tsurf[4];  // bottom, top, left, right (imc_mesh.jl:248-255)
  std::vector<N> fleck, beta, sigma_static, sa, sa_c, sa_p, ss, ss_c, ss_p, bee, radsource;
  std::vector<N> radenergydens, matenergydens, nrg_inc, energydep, emittedenergy;
  N totalenergy, totalenergydep, radenergyold;
  double lostenergy = 0; bool lost_wide = false;  // T; Float64 after an MC_RW vacuum loss of a Float64 energy (Q3)
  bool have_mesh = false;
  // particle list (Vector{Vector{T}}) + engine-side ids
  std::vector<Slots> particles;
  std::vector<uint64_t> ids;
  std::vector<int32_t> out_event, out_nseg;
  uint64_t iterations = 0;  // simvars.iterations, cumulative (Q13)
  // random-walk tables (RWVars)
  std::vector<N> aVals, prVals, ptVals;
  // tapes
  std::vector<double> tt_uni, tt_exp, st_uni; int tt_nuni = 0, tt_nexp = 0, st_nuni = 0; int64_t tt_slots = 0, st_slots = 0;
  // reduce buffer (single rank: plain copy of the tallies as doubles)
  std::vector<double> redbuf;

  explicit Oracle(const imc_config& c) : cfg(c) {
    geom = c.geometry; nx = c.nx; ny = geom == 2 ? c.ny : 1; nc = (size_t)nx * ny; ns = c.n_scales;
    nslots = geom == 1 ? 9 : 10;
    for (int k = 0; k < ns; ++k) scales.push_back(N::from_d(c.energyscales[k]));
    ds = N::from_d(c.distancescale); phys_c = N::from_d(c.phys_c); phys_a = N::from_d(c.phys_a); alpha = N::from_d(c.alpha);
  }

  size_t cidx(int xi, int yi) const { return (size_t)(xi - 1) + (size_t)nx * (yi - 1); }  // 1-based in
  double tempv(size_t i) const { return temp[i]; }

  int set_mesh(const double* dx_, const double* dy_, const double* sac, const double* sap, const double* ssc,
               const double* ssp, const double* sstat, const double* bee_, const double* rad, const double* temp_,
               const double* tsb, const double* tst, const double* tsl, const double* tsr) override {
    auto cp = [&](std::vector<N>& v, const double* src, size_t n) { v.resize(n); for (size_t i = 0; i < n; ++i) v[i] = src ? N::from_d(src[i]) : N(); };
    cp(dx, dx_, nx);
    if (geom == 2) cp(dy, dy_, ny); else dy.assign(1, N::from_d(1.0));
    cp(sa_c, sac, nc); cp(sa_p, sap, nc); cp(ss_c, ssc, nc); cp(ss_p, ssp, nc); cp(sigma_static, sstat, nc);
    sa = sa_c; ss = ss_c;  // column 1 starts as a copy of column 2 (imc_mesh.jl:137-141)
    cp(bee, bee_, nc); cp(radsource, rad, nc);
    temp.resize(nc); for (size_t i = 0; i < nc; ++i) temp[i] = N::from_d(temp_[i]).d();
    temp_wide = false;
    if (geom == 1) { tsurf1d[0] = N::from_d(tsl[0]); tsurf1d[1] = N::from_d(tsr[0]); }
    else { cp(tsurf[0], tsb, nx); cp(tsurf[1], tst, nx); cp(tsurf[2], tsl, ny); cp(tsurf[3], tsr, ny); }
    fleck.assign(nc, N()); beta.assign(nc, N::from_d(1.0));
    radenergydens.assign(nc, N()); matenergydens.assign(nc, N()); nrg_inc.assign(nc, N());
    energydep.assign(nc * ns, N()); emittedenergy.assign(nc * ns, N());
    totalenergy = N(); totalenergydep = N(); radenergyold = N(); lostenergy = 0; lost_wide = false;
    have_mesh = true;
    return IMC_OK;
  }

  // ---- Transport.P_r (imc_transport.jl:734-754): Float64 accumulation (Q31) --------------
  double P_r(double a, bool* is_T_one) const {
    *is_T_one = false;
    if (a != 0) {
      double Pr = 0.0;
      for (int n = 1; n <= 100; ++n) {
        double pin = 3.141592653589793 * (double)n;
        double sgn = ((n - 1) & 1) ? -1.0 : 1.0;
        Pr += sgn * M::exp64(-a * (pin * pin)) * 2.0;
      }
      return Pr;
    }
    *is_T_one = true;
    return 1.0;
  }
  // Transport.bisection (imc_transport.jl:756-784), 1-based result
  static int bisection(const std::vector<N>& arr, double value) {
    int n = (int)arr.size();
    if (value < arr[0].d()) return 1;
    if (value > arr[n - 1].d()) return n;
    int jl = 1, ju = n;
    while (ju - jl > 1) {
      int jm = (ju + jl) >> 1;
      if (value >= arr[jm - 1].d()) jl = jm; else ju = jm;
    }
    if (value == arr[0].d()) return 1;
    if (value == arr[n - 1].d()) return n;
    return jl;
  }
  // Transport.randomwalk_table (imc_transport.jl:786-797) with aVals = T.(LinRange(lo, hi, n))
  int rw_table(double lo, double hi, int n, double* a_out, double* pr_out, double* pt_out) override {
    aVals.resize(n); prVals.resize(n); ptVals.resize(n);
    for (int i = 0; i < n; ++i) {
      // LinRange element i (0-based): lerp in Float64 as Base.lerpi does: (1-t)*a + t*b, t = i/(n-1)
      double t = n > 1 ? (double)i / (double)(n - 1) : 0.0;
      double a = (1.0 - t) * lo + t * hi;
      aVals[i] = N::from_d(a);
      bool one;
      double pr = P_r(aVals[i].d(), &one);
      prVals[i] = N::from_d(pr);           // stored into zeros(T)
      ptVals[i] = N::from_d(1.0) - prVals[i];  // T(1 - prVals[i]) : Int - T is T
      if (a_out) a_out[i] = aVals[i].d();
      if (pr_out) pr_out[i] = prVals[i].d();
      if (pt_out) pt_out[i] = ptVals[i].d();
    }
    return IMC_OK;
  }

  // ---- Update.update (imc_update.jl:12-70) ------------------------------------------------
  int update(double dt_) override {
    if (!have_mesh) { err = "update before set_mesh"; return IMC_ERR_STATE; }
    N dt = N::from_d(dt_);
    N fourA = N::from_i(4) * phys_a;  // 4 * phys_a in T
    for (size_t i = 0; i < nc; ++i) {
      double t = temp[i];
      // 4 * phys_a * temp^3 : temp^3 is literal_pow -> (t*t)*t in temp's element type
      N four_a_t3; double four_a_t3_w = 0;
      if (temp_wide) four_a_t3_w = fourA.d() * ((t * t) * t);
      else { N tt = N::from_d(t); four_a_t3 = fourA * ((tt * tt) * tt); }
      if (cfg.linearized) {
        bee[i] = temp_wide ? N::from_d(four_a_t3_w) : four_a_t3;  // bee .= ... (stored in T)  :24
        beta[i] = N::from_d(1.0);                                  // :25
      } else {
        beta[i] = temp_wide ? N::from_d(four_a_t3_w / bee[i].d()) : four_a_t3 / bee[i];  // :27
      }
      // opacities :31-35 / :46-47
      if (temp_wide) {
        sa[i] = N::from_d(sa_c[i].d() * M::pow64(t, sa_p[i].d()));
        if (geom == 1 && cfg.marshak_quirk) sa[i] = N::from_d(((sa_c[i].d() / t) / t) / t);  // Q18
        ss[i] = N::from_d(ss_c[i].d() * M::pow64(t, ss_p[i].d()));
      } else {
        N tt = N::from_d(t);
        sa[i] = sa_c[i] * M::template pow<P>(tt, sa_p[i]);
        if (geom == 1 && cfg.marshak_quirk) sa[i] = ((sa_c[i] / tt) / tt) / tt;
        ss[i] = ss_c[i] * M::template pow<P>(tt, ss_p[i]);
      }
    }
    N one = N::from_d(1.0);
    for (size_t i = 0; i < nc; ++i) {  // :38-43 / :50-54
      double vals[6] = {ds.d(), alpha.d(), beta[i].d(), phys_c.d(), dt.d(), sa[i].d()};
      SorterResult<P> r = sorter<P>(vals, 6, &one, 1);
      fleck[i] = N::from_d(1.0 / (1.0 + r.product.d()));  // Float64 arithmetic, then T(...)  (Q31)
    }
    return IMC_OK;
  }

  // ---- Sourcing.sourcing (imc_sourcing.jl:12-370) -----------------------------------------
  struct Draws {  // one new particle's random numbers: Philox or source tape
    Oracle* o; bool tape; PhiloxDraw<P> ph; TapeDraw<P> tp;
    N uniform() { return tape ? tp.uniform() : ph.uniform(); }
  };
  Draws source_draws(uint64_t id, int64_t step, int64_t ordinal) {
    Draws d; d.o = this; d.tape = cfg.rng_mode == IMC_RNG_TAPE;
    if (d.tape) d.tp.init(st_uni.data(), st_nuni, nullptr, 0, (size_t)st_slots, (size_t)ordinal);
    else d.ph.init((uint64_t)cfg.seed, id, (uint32_t)step, STREAM_SOURCE);
    return d;
  }
  bool tape_over = false;

  // e / n with n an integer count: T(e / T(n)); Float16 with counts beyond its range divides in Float32 (Q10)
  N div_count(N e, int64_t n) const {
    if constexpr (P::id == 0) { if (n > 65504) return N(P::rnd(e.v / (float)n)); }
    return e / N::from_i(n);
  }
  // tointeger(max(round(((e/escale)*n_source)/total), cellmin))  (imc_sourcing.jl:139) — count arithmetic in T,
  // or in Float32 for Float16 decks whose counts exceed Float16 (Q10, intentional divergence)
  int64_t count_of(N e, N escale, double nsrc, N total, N cellmin, bool floor_cellmin, bool wide_counts, bool* bad) const {
    double r;
    if (wide_counts) {
      using W = Num<F32>;
      W x = ((W(e.v) / W(escale.v)) * W::from_d(nsrc)) / W(total.v);
      x = jl_round(x);
      if (floor_cellmin) x = jl_max(x, W(cellmin.v));
      r = x.d();
    } else {
      N x = ((e / escale) * N::from_d(nsrc)) / total;
      x = jl_round(x);
      if (floor_cellmin) x = jl_max(x, cellmin);
      r = x.d();
    }
    if (!(r - r == 0.0) || r < 0) { *bad = true; return 0; }  // InexactError in the reference
    return (int64_t)r;
  }

  int source(double dt_, int64_t n_input, double cellmin_, int64_t step, int64_t n_census_global, imc_source_stats* out) override {
    if (!have_mesh) { err = "source before set_mesh"; return IMC_ERR_STATE; }
    if (step < 0 || step >= (1ll << 24)) { err = "time-step index outside [0, 2^24)"; return IMC_ERR_ARG; }
    N dt = N::from_d(dt_), cellmin = N::from_d(cellmin_);
    const N* sc = scales.data();
    std::vector<N> e_body(nc), e_rad(nc), es_body(nc, N::from_d(1.0)), es_rad(nc, N::from_d(1.0)), es_em(nc, N::from_d(1.0));
    std::vector<N> e_sb(nx), e_st(nx), e_sl(ny), e_sr(ny), es_sb(nx, N::from_d(1.0)), es_st(nx, N::from_d(1.0)), es_sl(ny, N::from_d(1.0)), es_sr(ny, N::from_d(1.0));
    std::fill(emittedenergy.begin(), emittedenergy.end(), N());
    N e_surface, eL, sL, eR, sR;
    bool bad = false;
    if (geom == 1) {
      {  // :60-64
        double tl = tsurf1d[0].d(), tr = tsurf1d[1].d();
        double vl[8] = {phys_a.d(), phys_c.d(), tl, tl, tl, tl, dt.d(), 0.25};
        double vr[8] = {phys_a.d(), phys_c.d(), tr, tr, tr, tr, dt.d(), 0.25};
        auto rl = sorter<P>(vl, 8, sc, ns); auto rr = sorter<P>(vr, 8, sc, ns);
        eL = rl.product; sL = rl.scale; eR = rr.product; sR = rr.scale;
        e_surface = (eL / sL) + (eR / sR);
      }
      for (size_t i = 0; i < nc; ++i) {  // :66-71
        double t = temp[i];
        double vb[11] = {fleck[i].d(), sa[i].d(), phys_a.d(), phys_c.d(), t, t, t, t, dx[i].d(), dt.d(), ds.d()};
        auto rb = sorter<P>(vb, 11, sc, ns); e_body[i] = rb.product; es_body[i] = rb.scale;
        double vr[3] = {radsource[i].d(), dx[i].d(), dt.d()};
        auto rr = sorter<P>(vr, 3, sc, ns); e_rad[i] = rr.product; es_rad[i] = rr.scale;
        double ve[10] = {fleck[i].d(), sa[i].d(), phys_a.d(), phys_c.d(), t, t, t, t, dt.d(), ds.d()};
        auto re = sorter<P>(ve, 10, sc, ns); es_em[i] = re.scale;
        if (re.index >= 1) emittedenergy[i + nc * (re.index - 1)] = re.product; else bad = true;
      }
    } else {
      for (int i = 0; i < nx; ++i) {  // :86-89
        double tb = tsurf[0][i].d(), tt = tsurf[1][i].d();
        double vb[9] = {phys_a.d(), phys_c.d(), tb, tb, tb, tb, dx[i].d(), dt.d(), 0.25};
        double vt[9] = {phys_a.d(), phys_c.d(), tt, tt, tt, tt, dx[i].d(), dt.d(), 0.25};
        auto rb = sorter<P>(vb, 9, sc, ns); e_sb[i] = rb.product; es_sb[i] = rb.scale;
        auto rt = sorter<P>(vt, 9, sc, ns); e_st[i] = rt.product; es_st[i] = rt.scale;
      }
      for (int j = 0; j < ny; ++j) {  // :90-93
        double tl = tsurf[2][j].d(), tr = tsurf[3][j].d();
        double vl[9] = {phys_a.d(), phys_c.d(), tl, tl, tl, tl, dy[j].d(), dt.d(), 0.25};
        double vr[9] = {phys_a.d(), phys_c.d(), tr, tr, tr, tr, dy[j].d(), dt.d(), 0.25};
        auto rl = sorter<P>(vl, 9, sc, ns); e_sl[j] = rl.product; es_sl[j] = rl.scale;
        auto rr = sorter<P>(vr, 9, sc, ns); e_sr[j] = rr.product; es_sr[j] = rr.scale;
      }
      auto ratio_sum = [&](const std::vector<N>& e, const std::vector<N>& s) { std::vector<N> q(e.size()); for (size_t i = 0; i < e.size(); ++i) q[i] = e[i] / s[i]; return jl_sum(q); };
      e_surface = ((ratio_sum(e_sb, es_sb) + ratio_sum(e_st, es_st)) + ratio_sum(e_sl, es_sl)) + ratio_sum(e_sr, es_sr);  // :95
      for (int yi = 1; yi <= ny; ++yi) for (int xi = 1; xi <= nx; ++xi) {  // :97-107 (CartesianIndices: x fastest)
        size_t i = cidx(xi, yi);
        double t = temp[i];
        double vb[12] = {fleck[i].d(), sa[i].d(), phys_a.d(), phys_c.d(), t, t, t, t, dx[xi - 1].d(), dy[yi - 1].d(), dt.d(), ds.d()};
        auto rb = sorter<P>(vb, 12, sc, ns); e_body[i] = rb.product; es_body[i] = rb.scale;
        double vr[4] = {radsource[i].d(), dx[xi - 1].d(), dy[yi - 1].d(), dt.d()};
        auto rr = sorter<P>(vr, 4, sc, ns); e_rad[i] = rr.product; es_rad[i] = rr.scale;
        double ve[10] = {fleck[i].d(), sa[i].d(), phys_a.d(), phys_c.d(), t, t, t, t, dt.d(), ds.d()};
        auto re = sorter<P>(ve, 10, sc, ns); es_em[i] = re.scale;
        if (re.index >= 1) emittedenergy[i + nc * (re.index - 1)] = re.product; else bad = true;
      }
    }
    {  // :121
      std::vector<N> qb(nc), qr(nc);
      for (size_t i = 0; i < nc; ++i) { qb[i] = e_body[i] / es_body[i]; qr[i] = e_rad[i] / es_rad[i]; }
      totalenergy = (jl_sum(qb) + jl_sum(qr)) + e_surface;
    }
    double emitted_sum = 0;
    {  // print at :75 : sum(mesh.emittedenergy ./ escale_emittedenergy) — broadcast of (Nc x Ns) ./ (Nc)
      std::vector<N> q(nc * ns);
      for (int k = 0; k < ns; ++k) for (size_t i = 0; i < nc; ++i) q[i + nc * k] = emittedenergy[i + nc * k] / es_em[i];
      emitted_sum = jl_sum(q).d();
    }
    // number of particles to source :132-136 (Q9)
    int64_t n_census = n_census_global >= 0 ? n_census_global : (int64_t)particles.size();
    // Float16 decks whose counts exceed Float16's range do the count arithmetic in Float32 (Q10)
    bool wide_counts = (P::id == 0) && (std::max<int64_t>(n_input, cfg.n_max) > 65504);
    auto toT = [&](int64_t v) { return wide_counts ? (double)(float)v : N::from_i(v).d(); };
    double nsrc = toT(n_input);
    if (n_input + n_census > cfg.n_max) {
      int64_t cand = cfg.n_max - n_census - (geom == 1 ? 1 : 2) - 1;  // length(Ncells) is 1 or 2 (Q9)
      double cand_t = toT(cand);                                      // max(cellmin::T, ::Int) promotes to T
      nsrc = cellmin.d() > cand_t ? cellmin.d() : cand_t;
    }
    int64_t n_source_i = (int64_t)nsrc;

    std::vector<int64_t> n_body(nc), n_rad(nc, 0);
    for (size_t i = 0; i < nc; ++i) n_body[i] = count_of(e_body[i], es_body[i], nsrc, totalenergy, cellmin, true, wide_counts, &bad);  // :138-140
    for (size_t i = 0; i < nc; ++i) if (e_rad[i] > N()) n_rad[i] = count_of(e_rad[i], es_rad[i], nsrc, totalenergy, cellmin, true, wide_counts, &bad);  // :142-146

    const int64_t world = cfg.world > 0 ? cfg.world : 1, rank = cfg.rank;
    int64_t ordinal = 0, n_new_local = 0;
    auto mine = [&](int64_t j) { return j % world == rank; };
    auto new_id = [&](int64_t j) { return ((uint64_t)step << 40) | (uint64_t)j; };
    N one = N::from_d(1.0);
    if (geom == 1) {
      int64_t n_left = 0, n_right = 0;  // :150-157 (no cellmin floor, Q30)
      if (eL > N()) n_left = count_of(eL, sL, nsrc, totalenergy, cellmin, false, wide_counts, &bad);
      if (eR > N()) n_right = count_of(eR, sR, nsrc, totalenergy, cellmin, false, wide_counts, &bad);
      for (int64_t q = 0; q < n_left; ++q, ++ordinal) {  // :161-174
        if (!mine(ordinal)) continue;
        Draws d = source_draws(new_id(ordinal), step, ordinal);
        N origin = N::from_i(1);
        N xpos = N::from_d((0.01 * dx[0].d()) * ds.d());
        N nrg = div_count(eL, n_left);
        N mu = M::template sqrt<P>(d.uniform());
        while (mu == N()) mu = M::template sqrt<P>(d.uniform());
        N spawn = dt * d.uniform();
        push1d(origin, spawn, origin, xpos, mu, nrg, sL, new_id(ordinal)); ++n_new_local;
        tape_over |= d.tape && d.tp.exhausted();
      }
      for (int64_t q = 0; q < n_right; ++q, ++ordinal) {  // :176-189
        if (!mine(ordinal)) continue;
        Draws d = source_draws(new_id(ordinal), step, ordinal);
        N origin = N::from_i((long long)nc);
        N xpos = N::from_d((0.99 * dx[nc - 1].d()) * ds.d());
        N nrg = div_count(eR, n_right);
        N mu = -M::template sqrt<P>(d.uniform());
        while (mu == N()) mu = -M::template sqrt<P>(d.uniform());
        N spawn = dt * d.uniform();
        push1d(origin, spawn, origin, xpos, mu, nrg, sR, new_id(ordinal)); ++n_new_local;
        tape_over |= d.tape && d.tp.exhausted();
      }
      for (int pass = 0; pass < 2; ++pass) {  // body :193-214, then radiation source :217-236
        const std::vector<int64_t>& cnt = pass == 0 ? n_body : n_rad;
        const std::vector<N>& en = pass == 0 ? e_body : e_rad;
        const std::vector<N>& es = pass == 0 ? es_body : es_rad;
        for (size_t c = 0; c < nc; ++c) {
          if (cnt[c] <= 0) continue;
          N nrg = div_count(en[c], cnt[c]);
          for (int64_t q = 0; q < cnt[c]; ++q, ++ordinal) {
            if (!mine(ordinal)) continue;
            Draws d = source_draws(new_id(ordinal), step, ordinal);
            N cell = N::from_i((long long)c + 1);
            N xpos = (dx[c] * d.uniform()) * ds;
            N mu = one - N::from_i(2) * d.uniform();
            while (mu == N()) mu = one - N::from_i(2) * d.uniform();
            N spawn = dt * d.uniform();
            push1d(cell, spawn, cell, xpos, mu, nrg, es[c], new_id(ordinal)); ++n_new_local;
            tape_over |= d.tape && d.tp.exhausted();
          }
        }
      }
    } else {
      // 2-D surface counts :240-263 (with cellmin floor)
      std::vector<int64_t> n_sb(nx, 0), n_st(nx, 0), n_sl(ny, 0), n_sr(ny, 0);
      for (int i = 0; i < nx; ++i) if (e_sb[i] > N()) n_sb[i] = count_of(e_sb[i], es_sb[i], nsrc, totalenergy, cellmin, true, wide_counts, &bad);
      for (int i = 0; i < nx; ++i) if (e_st[i] > N()) n_st[i] = count_of(e_st[i], es_st[i], nsrc, totalenergy, cellmin, true, wide_counts, &bad);
      for (int j = 0; j < ny; ++j) if (e_sl[j] > N()) n_sl[j] = count_of(e_sl[j], es_sl[j], nsrc, totalenergy, cellmin, true, wide_counts, &bad);
      for (int j = 0; j < ny; ++j) if (e_sr[j] > N()) n_sr[j] = count_of(e_sr[j], es_sr[j], nsrc, totalenergy, cellmin, true, wide_counts, &bad);
      const double PI = 3.141592653589793;
      N pi_T = N::from_d(PI);
      for (int side = 0; side < 4; ++side) {  // bottom :265-278, top :280-293, left :295-308, right :310-323
        const std::vector<int64_t>& cnt = side == 0 ? n_sb : side == 1 ? n_st : side == 2 ? n_sl : n_sr;
        const std::vector<N>& en = side == 0 ? e_sb : side == 1 ? e_st : side == 2 ? e_sl : e_sr;
        const std::vector<N>& es = side == 0 ? es_sb : side == 1 ? es_st : side == 2 ? es_sl : es_sr;
        int len = side < 2 ? nx : ny;
        for (int i = 0; i < len; ++i) {
          for (int64_t q = 0; q < cnt[i]; ++q, ++ordinal) {
            if (!mine(ordinal)) continue;
            Draws d = source_draws(new_id(ordinal), step, ordinal);
            N spawn = dt * d.uniform();
            N xi, yi, xpos, ypos, mu;
            if (side == 0) {
              xi = N::from_i(i + 1); yi = N::from_i(1);
              xpos = (dx[i] * d.uniform()) * ds;
              ypos = N::from_d((0.001 * dy[0].d()) * ds.d());
              mu = pi_T * d.uniform();                       // precision(pi*rand(T)) : Irrational*T in T
            } else if (side == 1) {
              xi = N::from_i(i + 1); yi = N::from_i(ny);
              xpos = (dx[i] * d.uniform()) * ds;
              ypos = N::from_d((0.999 * dy[ny - 1].d()) * ds.d());
              mu = N::from_d((-PI) * d.uniform().d());         // -pi is Float64  (Q31)
            } else if (side == 2) {
              xi = N::from_i(1); yi = N::from_i(i + 1);
              xpos = N::from_d((0.001 * dx[0].d()) * ds.d());
              ypos = (dx_q7(i) * d.uniform()) * ds;            // Q7: mesh.dx[j]
              mu = N::from_d(PI * (0.5 - d.uniform().d()));
            } else {
              xi = N::from_i(nx); yi = N::from_i(i + 1);
              xpos = N::from_d((0.999 * dx[nx - 1].d()) * ds.d());
              ypos = (dx_q7(i) * d.uniform()) * ds;
              mu = N::from_d(PI * (0.5 + d.uniform().d()));
            }
            N nrg = div_count(en[i], cnt[i]);                 // e / n (n stored in zeros(T))
            push2d(spawn, xi, yi, xpos, ypos, mu, nrg, es[i], new_id(ordinal)); ++n_new_local;
            tape_over |= d.tape && d.tp.exhausted();
          }
        }
      }
      for (int pass = 0; pass < 2; ++pass) {  // body :326-346, radiation source :349-366
        const std::vector<N>& en = pass == 0 ? e_body : e_rad;
        const std::vector<N>& es = pass == 0 ? es_body : es_rad;
        for (int yi = 1; yi <= ny; ++yi) for (int xi = 1; xi <= nx; ++xi) {
          size_t c = cidx(xi, yi);
          int64_t gate = pass == 0 ? n_body[c] : n_rad[c];
          if (gate <= 0) continue;
          N nrg = div_count(en[c], gate);
          int64_t loops = n_body[c];  // Q6: the radiation-source loop also runs 1:n_body
          for (int64_t q = 0; q < loops; ++q, ++ordinal) {
            if (!mine(ordinal)) continue;
            Draws d = source_draws(new_id(ordinal), step, ordinal);
            N xpos = (dx[xi - 1] * d.uniform()) * ds;
            N ypos = (dy[yi - 1] * d.uniform()) * ds;
            N mu = N::from_d((2.0 * PI) * d.uniform().d());
            N spawn = dt * d.uniform();
            push2d(spawn, N::from_i(xi), N::from_i(yi), xpos, ypos, mu, nrg, es[c], new_id(ordinal)); ++n_new_local;
            tape_over |= d.tape && d.tp.exhausted();
          }
        }
      }
    }
    if (out) {
      out->totalenergy = totalenergy.d(); out->emitted_sum = emitted_sum; out->n_source = n_source_i;
      out->n_new_global = ordinal; out->n_new_local = n_new_local; out->n_particles = (int64_t)particles.size();
    }
    if (tape_over) { err = "source tape exhausted"; return IMC_ERR_TAPE; }
    if (bad) { err = "non-finite particle count or unrepresentable energy (reference would throw)"; return IMC_ERR_NUMERIC; }
    return IMC_OK;
  }
  N dx_q7(int j) const { return j < nx ? dx[j] : dy[j]; }  // reference indexes dx with the y index (BoundsError if j > Nx)
  void push1d(N origin, N t, N cell, N x, N mu, N nrg, N escale, uint64_t id) {
    Slots s{}; s[0] = origin; s[1] = t; s[2] = cell; s[3] = x; s[4] = mu; s[5] = N::from_d(1.0); s[6] = nrg; s[7] = nrg; s[8] = escale;
    particles.push_back(s); ids.push_back(id);
  }
  void push2d(N t, N xi, N yi, N x, N y, N mu, N nrg, N escale, uint64_t id) {
    Slots s{}; s[0] = t; s[1] = xi; s[2] = yi; s[3] = x; s[4] = y; s[5] = mu; s[6] = N::from_d(1.0); s[7] = nrg; s[8] = nrg; s[9] = escale;
    particles.push_back(s); ids.push_back(id);
  }

  // ---- tally containers shared by MC / MC_RW / MC2D ----------------------------------------
  struct Dep {
    Oracle* o; bool pairwise;
    std::vector<std::vector<N>> dep_vec, lost_vec;
    void begin() {
      pairwise = o->cfg.pairwise != 0;
      std::fill(o->energydep.begin(), o->energydep.end(), N());
      if (pairwise) { dep_vec.assign(o->nc * o->ns, {}); lost_vec.assign(o->ns, {}); }
    }
    void add(size_t cell, int k, N v) {  // k 0-based scale plane
      if (pairwise) dep_vec[cell + o->nc * k].push_back(v); else o->energydep[cell + o->nc * k] += v;
    }
    void add_wide(size_t cell, int k, double v) {  // MC_RW: Float64 deposit into a T container (Q2/Q3)
      if (pairwise) dep_vec[cell + o->nc * k].push_back(N::from_d(v));
      else o->energydep[cell + o->nc * k] = N::from_d(o->energydep[cell + o->nc * k].d() + v);
    }
    void lose(int k, N energy, N scale) {
      if (pairwise) lost_vec[k].push_back(energy);
      else if (o->lost_wide) o->lostenergy = o->lostenergy + (energy / scale).d();
      else o->lostenergy = (N::from_d(o->lostenergy) + energy / scale).d();
    }
    void lose_wide(int k, double energy, N scale) {  // Float64 energy: mesh.lostenergy becomes Float64
      if (pairwise) lost_vec[k].push_back(N::from_d(energy));
      else { o->lostenergy = o->lostenergy + energy / scale.d(); o->lost_wide = true; }
    }
    void end() {
      if (!pairwise) return;
      for (int k = 0; k < o->ns; ++k) {
        N add = jl_sum(lost_vec[k]) / o->scales[k];
        o->lostenergy = o->lost_wide ? o->lostenergy + add.d() : (N::from_d(o->lostenergy) + add).d();
        for (size_t i = 0; i < o->nc; ++i) o->energydep[i + o->nc * k] = jl_sum(dep_vec[i + o->nc * k]);
      }
    }
  };

  int scale_index(N escale) const {  // findfirst(isequal(scale), energyscales), 0-based; -1 if absent
    for (int k = 0; k < ns; ++k) if (scales[k] == escale) return k;
    return -1;
  }

  // MC / MC2D: per-segment reserved Philox words (SegDraw); MC_RW: sequential Philox stream; replay: tape
  struct TDraws {
    bool tape, seg; PhiloxDraw<P> ph; SegDraw<P> sg; TapeDraw<P> tp; uint64_t seed; uint32_t step;
    void next_segment() { if (!tape && seg) sg.next_segment(seed, step); }
    N uniform() { return tape ? tp.uniform() : seg ? sg.uniform(seed, step) : ph.uniform(); }
    N randexp() { return tape ? tp.randexp() : seg ? sg.randexp() : ph.randexp(); }
    double randexp64() { return tape ? tp.randexp64() : ph.randexp64(); }
    bool over() const { return tape && tp.exhausted(); }
  };
  TDraws track_draws(size_t slot, int64_t step, bool per_segment = true) {
    TDraws d; d.tape = cfg.rng_mode == IMC_RNG_TAPE; d.seg = per_segment; d.seed = (uint64_t)cfg.seed; d.step = (uint32_t)step;
    if (d.tape) d.tp.init(tt_uni.data(), tt_nuni, tt_exp.data(), tt_nexp, (size_t)tt_slots, slot);
    else if (per_segment) d.sg.init(ids[slot]);
    else d.ph.init((uint64_t)cfg.seed, ids[slot], (uint32_t)step, STREAM_TRACK);
    return d;
  }

  int transport(double dt_, int64_t step, imc_transport_stats* out) override {
    if (!have_mesh) { err = "transport before set_mesh"; return IMC_ERR_STATE; }
    if (step < 0 || step >= (1ll << 24)) { err = "time-step index outside [0, 2^24)"; return IMC_ERR_ARG; }
    if (cfg.rng_mode == IMC_RNG_TAPE && (int64_t)particles.size() > tt_slots) { err = "transport tape has fewer slots than particles"; return IMC_ERR_TAPE; }
    imc_transport_stats st{};
    out_event.assign(particles.size(), 0); out_nseg.assign(particles.size(), 0);
    uint64_t before = iterations;
    int rc;
    if (geom == 1) rc = cfg.randomwalk ? MC_RW(N::from_d(dt_), step, st) : MC(N::from_d(dt_), step, st);
    else rc = MC2D(N::from_d(dt_), step, st);
    st.segments = iterations - before; st.segments_total = iterations; st.histories = (int64_t)particles.size();
    st.lostenergy = lostenergy; st.variant = IMC_TRACK_HISTORY; st.tally_mode = IMC_TALLY_EXACT; st.kernel_ms = 0;
    if (out) *out = st;
    return rc;
  }

  // ---- Transport.MC (imc_transport.jl:13-210) ----------------------------------------------
  int MC(N dt, int64_t step, imc_transport_stats& st) {
    Dep dep{this}; dep.begin();
    bool over = false;
    N one = N::from_d(1.0);
    for (size_t p = 0; p < particles.size(); ++p) {
      Slots& s = particles[p];
      N origin = s[0], t = s[1]; long long cell = (long long)s[2].d(); N x = s[3], mu = s[4], freq = s[5], E = s[6], E0 = s[7], escale = s[8];
      int k = scale_index(escale);
      if (k < 0 || cell < 1 || cell > (long long)nc) { ++st.n_errors; s[7] = N::from_d(-1.0); out_event[p] = 1; continue; }
      N minenergy = N::from_d(0.01 * E0.d());  // :61 (Float64 product, Q31)
      TDraws d = track_draws(p, step);
      int nseg = 0, ev = 0;
      while (true) {
        ++iterations; ++nseg;  // :73
        d.next_segment();
        size_t c = (size_t)cell - 1;
        N dist_b = mu > N() ? (dx[c] * ds - x) / mu : nabs(x / mu);               // :77-83
        N dist_col = d.randexp() / (sa[c] * (one - fleck[c]) + ss[c]);             // :87
        N dist_cen = (phys_c * (dt - t)) * ds;                                     // :89
        N dist = jl_min(jl_min(dist_b, dist_col), dist_cen);                       // :92
        N arg = ((-sa[c]) * fleck[c]) * dist;
        N ex, em1; M::template exp_expm1<P>(arg, &ex, &em1);
        N newE = E * ex;                                                           // :95
        if (is_nan(newE) || is_nan(dist)) ++st.n_errors;
        if (newE <= minenergy) {                                                   // :97-106
          dep.add(c, k, E / dx[c]);
          s[7] = N::from_d(-1.0); ev = 1; ++st.n_absorbed;
          break;
        }
        dep.add(c, k, (-(E / dx[c])) * em1);                                       // :110 / :120
        x = x + mu * dist;                                                         // :124
        t = t + (dist / ds) / phys_c;                                              // :125
        E = newE;                                                                  // :126
        bool dead = false;
        if (dist == dist_b) {                                                      // :130-170
          if (mu > N()) {
            if (cell == (long long)nc) {
              if (cfg.bc[IMC_BC_RIGHT] == IMC_REFLECT) mu = -mu;
              else { dep.lose(k, E, escale); s[7] = N::from_d(-1.0); dead = true; }
            }
            if (!dead) { cell += 1; x = N(); }
          }
          if (!dead && mu < N()) {
            if (cell == 1) {
              if (cfg.bc[IMC_BC_LEFT] == IMC_REFLECT) mu = -mu;
              else { dep.lose(k, E, escale); s[7] = N::from_d(-1.0); dead = true; }
            } else { cell -= 1; x = dx[cell - 1] * ds; }
          }
        }
        if (dead) { ev = 2; ++st.n_escaped; break; }
        if (dist == dist_col) {                                                    // :174-183
          mu = N();
          while (mu == N()) mu = one - N::from_i(2) * d.uniform();
        }
        if (dist == dist_cen) {                                                    // :185-193
          t = N();
          s[0] = origin; s[1] = t; s[2] = N::from_i(cell); s[3] = x; s[4] = mu; s[5] = freq; s[6] = E; s[7] = E0; s[8] = escale;
          ev = 0; ++st.n_census;
          break;
        }
      }
      out_event[p] = ev; out_nseg[p] = nseg;
      over |= d.over();
    }
    dep.end();
    if (over) { err = "transport tape exhausted"; return IMC_ERR_TAPE; }
    return IMC_OK;
  }

  // ---- Transport.MC_RW (imc_transport.jl:212-479) ------------------------------------------
  int MC_RW(N dt, int64_t step, imc_transport_stats& st) {
    if (aVals.empty()) { err = "random-walk tables not set (imc_rw_table)"; return IMC_ERR_STATE; }
    Dep dep{this}; dep.begin();
    bool over = false;
    N one = N::from_d(1.0);
    using D = Dyn<P>;
    for (size_t p = 0; p < particles.size(); ++p) {
      Slots& s = particles[p];
      N origin = s[0]; D t(s[1]); long long cell = (long long)s[2].d(); D x(s[3]); N mu = s[4], freq = s[5]; D E(s[6]); N E0 = s[7], escale = s[8];
      int k = scale_index(escale);
      if (k < 0 || cell < 1 || cell > (long long)nc) { ++st.n_errors; s[7] = N::from_d(-1.0); out_event[p] = 1; continue; }
      N minenergy = N::from_d(0.01 * E0.d());
      TDraws d = track_draws(p, step, false);
      int nseg = 0, ev = 0;
      while (true) {
        ++iterations; ++nseg;                                                      // :265
        size_t c = (size_t)cell - 1;
        D dist_b = mu > N() ? (D(dx[c]) - x) / D(mu) : dabs(x / D(mu));            // :269-275 (no distancescale, Q3)
        D dist_col = D::w(std::fabs(d.randexp64()) / (sa[c] * (one - fleck[c]) + ss[c]).d());  // :279 Float64
        D dist_cen = D(phys_c) * (D(dt) - t);                                      // :281
        D dist = dmin(dmin(dist_b, dist_col), dist_cen);                           // :284 (Float64 after promotion)
        D R0 = dmin(dabs(D(dx[c]) - x), dabs(x));                                  // :286
        N inv_sigma = N::from_i(1) / sigma_static[c];                              // 1/mesh.sigma[cellindex]  (Q4)
        if (R0.v > inv_sigma.d() && dist_col.v < R0.v) {                           // :289
          ++st.n_rw;
          N u = d.uniform();                                                       // :290
          N Dc = phys_c / ((N::from_i(3) * sa[c]) * (one - fleck[c]));             // :292
          D a = (D(Dc) * D(dt)) / (R0 * R0);                                       // :294
          bool one_T; double Pr = P_r(a.v, &one_T);                                // :296
          double Pt = 1.0 - Pr;                                                    // :297
          N lg = M::template log<P>(one - fleck[c]);
          N expo;
          if (u.d() < Pt) {                                                        // :298
            int a_index = bisection(ptVals, u.d());                                // :301
            D tp_d = (D(aVals[a_index - 1]) * (R0 * R0)) / D(Dc);
            N t_p = N::from_d(tp_d.v);                                             // :303
            expo = (((t_p * phys_c) * (one - fleck[c])) * sa[c]) / lg;             // :306
          } else {
            (void)d.uniform();                                                     // u_prime :338 (consumed; R1 unused)
            expo = ((((phys_c * (one - fleck[c])) * sa[c]) * dt) / lg);            // :346
          }
          // newenergy = energy * exp(expo); `if newenergy <= startenergy` is always true (Q1) -> particle dies
          D newE = E * D(M::template exp<P>(expo));
          D depv = (D(-one) * (E / D(dx[c]))) * D(M::template expm1<P>(expo));     // -(energy/dx)*expm1(...)  :317-320 / :352-356
          if (depv.wide) dep.add_wide(c, k, depv.v); else dep.add(c, k, depv.narrow());
          if (newE.v != newE.v) ++st.n_errors;  // NaN: the reference falls through into unreachable-by-design code
          s[7] = N::from_d(-1.0); ev = 3; ++st.n_absorbed;
          break;
        }
        // normal path :376-399
        D newE = E * D::w(M::exp64(((-sa[c]) * fleck[c]).d() * dist.v));           // Float64: dist is Float64
        if (newE.v <= minenergy.d()) newE = D(N());                                // :377-379
        D depv = E - newE;                                                         // :383 / :385 (not / dx, Q2)
        if (depv.wide) dep.add_wide(c, k, depv.v); else dep.add(c, k, depv.narrow());
        if (newE.v == 0.0) { s[7] = N::from_d(-1.0); ev = 1; ++st.n_absorbed; break; }  // :390-394
        x = x + D(mu) * dist;                                                      // :397
        t = t + dist / D(phys_c);                                                  // :398
        E = newE;                                                                  // :399
        bool dead = false;
        if (dist.v == dist_b.v) {                                                  // :403-443
          if (mu > N()) {
            if (cell == (long long)nc) {
              if (cfg.bc[IMC_BC_RIGHT] == IMC_REFLECT) mu = -mu;
              else { if (E.wide) dep.lose_wide(k, E.v, escale); else dep.lose(k, E.narrow(), escale); s[7] = N::from_d(-1.0); dead = true; }
            }
            if (!dead) { cell += 1; x = D(N()); }
          }
          if (!dead && mu < N()) {
            if (cell == 1) {
              if (cfg.bc[IMC_BC_LEFT] == IMC_REFLECT) mu = -mu;
              else { if (E.wide) dep.lose_wide(k, E.v, escale); else dep.lose(k, E.narrow(), escale); s[7] = N::from_d(-1.0); dead = true; }
            } else { cell -= 1; x = D(dx[cell - 1]); }
          }
        }
        if (dead) { ev = 2; ++st.n_escaped; break; }
        if (dist.v == dist_col.v) {                                                // :446-453
          mu = one - N::from_i(2) * d.uniform();
          while (mu == N()) mu = one - N::from_i(2) * d.uniform();
        }
        if (dist.v == dist_cen.v) {                                                // :455-463
          s[0] = origin; s[1] = N(); s[2] = N::from_i(cell); s[3] = N::from_d(x.v); s[4] = mu; s[5] = freq; s[6] = N::from_d(E.v); s[7] = E0; s[8] = escale;
          ev = 0; ++st.n_census;
          break;
        }
      }
      out_event[p] = ev; out_nseg[p] = nseg;
      over |= d.over();
    }
    dep.end();
    if (over) { err = "transport tape exhausted"; return IMC_ERR_TAPE; }
    return IMC_OK;
  }

  // ---- Transport.MC2D (imc_transport.jl:483-732) -------------------------------------------
  int MC2D(N dt, int64_t step, imc_transport_stats& st) {
    Dep dep{this}; dep.begin();
    bool over = false;
    N one = N::from_d(1.0);
    const double TWO_PI = 2.0 * 3.141592653589793;
    for (size_t p = 0; p < particles.size(); ++p) {
      Slots& s = particles[p];
      N t = s[0]; long long xi = (long long)s[1].d(), yi = (long long)s[2].d(); N x = s[3], y = s[4], mu = s[5], frq = s[6], E = s[7], E0 = s[8], escale = s[9];
      int k = scale_index(escale);
      if (k < 0 || xi < 1 || xi > nx || yi < 1 || yi > ny) { ++st.n_errors; s[7] = N::from_d(-1.0); out_event[p] = 1; continue; }
      N minenergy = N::from_d(0.01 * E0.d());  // :531
      TDraws d = track_draws(p, step);
      int nseg = 0, ev = 0;
      while (true) {
        ++iterations; ++nseg;  // counted as in MC (the reference's MC2D has no counter, Q13)
        d.next_segment();
        N vx, vy; M::template sincos<P>(mu, &vy, &vx);                             // :534
        size_t c = cidx((int)xi, (int)yi);
        N dxc = dx[xi - 1], dyc = dy[yi - 1];
        N dist_bx = vx > N() ? nabs((dxc * ds - x) / vx) : nabs(x / vx);           // :538-542
        N dist_by = vy > N() ? nabs((dyc * ds - y) / vy) : nabs(y / vy);           // :544-548
        N dist_b = is_nan(dist_bx) ? dist_by : is_nan(dist_by) ? dist_bx : jl_min(dist_bx, dist_by);  // :551-557
        N dist_col = d.randexp() / (sa[c] * (one - fleck[c]) + ss[c]);             // :561
        N dist_cen = (phys_c * (dt - t)) * ds;                                     // :569
        N dist = jl_min(jl_min(dist_b, dist_col), dist_cen);                       // :571
        if (is_nan(dist) || dist_col < N()) ++st.n_errors;
        N arg = ((-fleck[c]) * sa[c]) * dist;
        N ex, em1; M::template exp_expm1<P>(arg, &ex, &em1);
        N newE = E * ex;                                                           // :580
        if (newE <= minenergy) {                                                   // :586-595
          dep.add(c, k, (E / dxc) / dyc);
          s[7] = N::from_d(-1.0); ev = 1; ++st.n_absorbed;
          break;
        }
        dep.add(c, k, ((-(E / dxc)) / dyc) * em1);                                 // :599 / :607
        x = x + dist * vx;                                                         // :615
        y = y + dist * vy;                                                         // :616
        t = t + (dist / ds) / phys_c;                                              // :617
        E = newE;                                                                  // :618
        if (dist == dist_bx || dist == dist_by) {                                  // :621
          bool dead = false;
          int side;  // which domain boundary, if any
          if (dist_bx < dist_by) {                                                 // :622
            if (vx > N()) {                                                        // cos(mu) > 0
              if (xi == nx) side = IMC_BC_RIGHT; else { side = -1; xi += 1; x = N(); }
            } else {
              if (xi == 1) side = IMC_BC_LEFT; else { side = -1; xi -= 1; x = dx[xi - 1] * ds; }
            }
            if (side >= 0) {
              if (cfg.bc[side] == IMC_REFLECT) mu = M::template atan2<P>(vy, -vx);  // xvec - [1,0]*(2 v'xvec) = (-vx, vy)  :626-628
              else { dep.lose(k, E, escale); s[7] = N::from_d(-1.0); dead = true; }
            }
          } else {
            if (vy > N()) {                                                        // sin(mu) > 0
              if (yi == ny) side = IMC_BC_TOP; else { side = -1; yi += 1; y = N(); }
            } else {
              if (yi == 1) side = IMC_BC_BOTTOM; else { side = -1; yi -= 1; y = dy[yi - 1] * ds; }
            }
            if (side >= 0) {
              if (cfg.bc[side] == IMC_REFLECT) mu = M::template atan2<P>(-vy, vx);  // (vx, -vy)  :666-668
              else { dep.lose(k, E, escale); s[7] = N::from_d(-1.0); dead = true; }
            }
          }
          if (dead) { ev = 2; ++st.n_escaped; break; }
          continue;                                                                // :703 (Q15)
        }
        if (dist == dist_col) mu = N::from_d(TWO_PI * d.uniform().d());            // :706-710
        if (dist == dist_cen) {                                                    // :712-717
          s[0] = N(); s[1] = N::from_i(xi); s[2] = N::from_i(yi); s[3] = x; s[4] = y; s[5] = mu; s[6] = frq; s[7] = E; s[8] = E0; s[9] = escale;
          ev = 0; ++st.n_census;
          break;
        }
      }
      out_event[p] = ev; out_nseg[p] = nseg;
      over |= d.over();
    }
    dep.end();
    if (over) { err = "transport tape exhausted"; return IMC_ERR_TAPE; }
    return IMC_OK;
  }

  // ---- Clean.clean (imc_clean.jl:6-19) ------------------------------------------------------
  int clean(int64_t* n_alive) override {
    // The reference walks the list backwards and deleteat!s every flagged particle, which is O(N * Ndead);
    // the result is the stable removal below (same surviving order), done in one O(N) pass so that the
    // CPU-baseline timing is not dominated by that quadratic loop.
    size_t w = 0;
    for (size_t i = 0; i < particles.size(); ++i) {
      if (particles[i][7].d() == -1.0) continue;
      if (w != i) { particles[w] = particles[i]; ids[w] = ids[i]; }
      ++w;
    }
    particles.resize(w); ids.resize(w);
    if (n_alive) *n_alive = (int64_t)particles.size();
    return IMC_OK;
  }

  // ---- Tally.tally (imc_tally.jl:11-149) ----------------------------------------------------
  int tally_local() override {  // census radiation energy density :81-113 (always per-cell vectors + sum, Q19)
    std::vector<std::vector<N>> vec(nc);
    for (size_t j = 0; j < particles.size(); ++j) {
      const Slots& s = particles[j];
      if (geom == 1) {
        size_t c = (size_t)s[2].d() - 1;
        vec[c].push_back(s[6] / (dx[c] * s[8]));                                   // :92
      } else {
        int xi = (int)s[1].d(), yi = (int)s[2].d();
        vec[cidx(xi, yi)].push_back(s[7] / ((dx[xi - 1] * dy[yi - 1]) * s[9]));    // :106
      }
    }
    for (size_t c = 0; c < nc; ++c) radenergydens[c] = jl_sum(vec[c]);
    if (cfg.world > 1) {  // sharded test runs: publish [energydep | radenergydens | lostenergy] for the host's all-reduce
      redbuf.assign(nc * ns + nc + 8, 0.0);
      for (size_t i = 0; i < nc * ns; ++i) redbuf[i] = energydep[i].d();
      for (size_t i = 0; i < nc; ++i) redbuf[nc * ns + i] = radenergydens[i].d();
      redbuf[nc * ns + nc] = lostenergy;
    }
    return IMC_OK;
  }
  int tally_finish(double t_, double dt_, imc_tally_stats* out) override {
    N dt = N::from_d(dt_);
    if (cfg.world > 1 && redbuf.size() == nc * ns + nc + 8) {  // summed over ranks by the host
      for (size_t i = 0; i < nc * ns; ++i) energydep[i] = N::from_d(redbuf[i]);
      for (size_t i = 0; i < nc; ++i) radenergydens[i] = N::from_d(redbuf[nc * ns + i]);
      lostenergy = N::from_d(redbuf[nc * ns + nc]).d();
    }
    N one = N::from_d(1.0);
    if (t_ == 0.0) {  // :29-32 (Q11)
      for (size_t i = 0; i < nc; ++i) {
        double t = temp[i];
        double v[10] = {fleck[i].d(), sa[i].d(), phys_a.d(), phys_c.d(), t, t, t, t, dt.d(), ds.d()};
        matenergydens[i] = sorter<P>(v, 10, &one, 1).product;
      }
    }
    totalenergydep = N();                                                          // :44
    std::fill(nrg_inc.begin(), nrg_inc.end(), N());
    for (int k = 0; k < ns; ++k) {                                                 // :47-57
      std::vector<N> q(nc);
      for (size_t i = 0; i < nc; ++i) {
        nrg_inc[i] = nrg_inc[i] + (energydep[i + nc * k] - emittedenergy[i + nc * k]) / scales[k];
        N vol = geom == 1 ? dx[i] : dx[i % nx] * dy[i / nx];
        q[i] = (energydep[i + nc * k] * vol) / scales[k];
      }
      totalenergydep = totalenergydep + jl_sum(q);
    }
    double max_temp = -INFINITY;
    for (size_t i = 0; i < nc; ++i) {
      matenergydens[i] = matenergydens[i] + nrg_inc[i];                            // :68
      if (cfg.linearized) temp[i] = M::pow64(matenergydens[i].d(), 0.25);          // :72 (Q12: Float64 from here on)
      else temp[i] = temp_wide ? temp[i] + (nrg_inc[i] / bee[i]).d() : (N::from_d(temp[i]) + nrg_inc[i] / bee[i]).d();  // :74
      if (temp[i] > max_temp || temp[i] != temp[i]) max_temp = temp[i];
    }
    if (cfg.linearized && P::id != 2) temp_wide = true;
    if (hist_cap > 0) {   // push!(mesh.temp_saved, copy(mesh.temp)) ... (:58, :138-142)
      if ((int64_t)hist_temp.size() == hist_cap) ++hist_dropped;
      else { hist_temp.push_back(temp); hist_mat.push_back(matenergydens); hist_rad.push_back(radenergydens); hist_inc.push_back(nrg_inc); }
    }
    if (out) {
      out->totalenergydep = totalenergydep.d();
      out->energy_increase = jl_sum(nrg_inc).d();
      out->max_temp = max_temp;
      std::vector<N> q(nc); for (size_t i = 0; i < nc; ++i) q[i] = matenergydens[i] + radenergydens[i];
      out->total_energy_density = jl_sum(q).d();
    }
    return IMC_OK;
  }

  // ---- EnergyCheck.energychecker (imc_energycheck.jl:19-37) ----------------------------------
  int energycheck(imc_energy_stats* out) override {
    std::vector<N> q(nc);
    for (size_t i = 0; i < nc; ++i) q[i] = geom == 1 ? radenergydens[i] * dx[i] : (radenergydens[i] * dx[i % nx]) * dy[i / nx];
    N radenergy = jl_sum(q);
    // (totalenergy - totalenergydep - (radenergy - radenergyold) - lostenergy) / totalenergy
    double change = (radenergy - radenergyold).d();
    double num;
    bool lost_T = !lost_wide;
    if (lost_T) num = (((totalenergy - totalenergydep) - (radenergy - radenergyold)) - N::from_d(lostenergy)).d();
    else num = ((totalenergy - totalenergydep) - (radenergy - radenergyold)).d() - lostenergy;
    double e = lost_T ? (N::from_d(num) / totalenergy).d() : num / totalenergy.d();
    if (out) { out->radenergy = radenergy.d(); out->radenergy_change = change; out->lostenergy = lostenergy; out->energy_error = e; }
    radenergyold = radenergy;
    lostenergy = 0; lost_wide = false;
    return IMC_OK;
  }

  int reduce_buffer(void** ptr, int64_t* n, int32_t* is_int) override {
    if (redbuf.size() != nc * ns + nc + 8) redbuf.assign(nc * ns + nc + 8, 0.0);
    *ptr = redbuf.data(); *n = (int64_t)redbuf.size(); *is_int = 0;
    return IMC_OK;
  }

  int get_field(int f, double* dst, int64_t n) override {
    auto put = [&](const std::vector<N>& v) { if ((int64_t)v.size() != n) return (int)IMC_ERR_ARG; for (size_t i = 0; i < v.size(); ++i) dst[i] = v[i].d(); return (int)IMC_OK; };
    switch (f) {
      case IMC_FIELD_TEMP: if ((int64_t)nc != n) return IMC_ERR_ARG; for (size_t i = 0; i < nc; ++i) dst[i] = temp[i]; return IMC_OK;
      case IMC_FIELD_FLECK: return put(fleck);
      case IMC_FIELD_BETA: return put(beta);
      case IMC_FIELD_BEE: return put(bee);
      case IMC_FIELD_SIGMA_A: return put(sa);
      case IMC_FIELD_SIGMA_S: return put(ss);
      case IMC_FIELD_ENERGYDEP: return put(energydep);
      case IMC_FIELD_EMITTEDENERGY: return put(emittedenergy);
      case IMC_FIELD_MATENERGYDENS: return put(matenergydens);
      case IMC_FIELD_RADENERGYDENS: return put(radenergydens);
      case IMC_FIELD_NRG_INC: return put(nrg_inc);
    }
    return IMC_ERR_ARG;
  }
  int set_state(const double* temp_, const double* mat, const double* rad) override {
    if (temp_) for (size_t i = 0; i < nc; ++i) temp[i] = temp_wide ? temp_[i] : N::from_d(temp_[i]).d();
    if (mat) for (size_t i = 0; i < nc; ++i) matenergydens[i] = N::from_d(mat[i]);
    if (rad) for (size_t i = 0; i < nc; ++i) radenergydens[i] = N::from_d(rad[i]);
    return IMC_OK;
  }
  // native-precision variants of the two calls above (include/imc.h): elements of type T, mesh.temp as Float64
  // once it has turned Float64 (Q12)
  int field_elsize(int f) override {
    if (f < 0 || f >= IMC_FIELD_COUNT_) return 0;
    return (f == IMC_FIELD_TEMP && temp_wide) ? 8 : P::bytes;
  }
  static void pack_native(double v, void* dst, size_t i, int elsize) {
    if (elsize == 8) static_cast<double*>(dst)[i] = v;
    else { typename P::store_t q = P::pack(P::from_d(v)); memcpy(static_cast<char*>(dst) + i * sizeof q, &q, sizeof q); }
  }
  static double unpack_native(const void* src, size_t i, int elsize) {
    if (elsize == 8) return static_cast<const double*>(src)[i];
    typename P::store_t q; memcpy(&q, static_cast<const char*>(src) + i * sizeof q, sizeof q); return (double)P::unpack(q);
  }
  int get_field_native(int f, void* dst, int64_t bytes) override {
    const int es = field_elsize(f);
    if (es == 0 || bytes % es != 0) return IMC_ERR_ARG;
    std::vector<double> tmp((size_t)(bytes / es));
    int rc = get_field(f, tmp.data(), (int64_t)tmp.size());
    if (rc) return rc;
    for (size_t i = 0; i < tmp.size(); ++i) pack_native(tmp[i], dst, i, es);
    return IMC_OK;
  }
  int set_state_native(const void* temp_, const void* mat, const void* rad) override {
    std::vector<double> a, b, c;
    if (temp_) { a.resize(nc); const int es = field_elsize(IMC_FIELD_TEMP); for (size_t i = 0; i < nc; ++i) a[i] = unpack_native(temp_, i, es); }
    if (mat) { b.resize(nc); for (size_t i = 0; i < nc; ++i) b[i] = unpack_native(mat, i, P::bytes); }
    if (rad) { c.resize(nc); for (size_t i = 0; i < nc; ++i) c[i] = unpack_native(rad, i, P::bytes); }
    return set_state(temp_ ? a.data() : nullptr, mat ? b.data() : nullptr, rad ? c.data() : nullptr);
  }
  // per-step history (include/imc.h): the reference's *_saved lists
  int64_t hist_cap = 0, hist_dropped = 0;
  std::vector<std::vector<double>> hist_temp;
  std::vector<std::vector<N>> hist_mat, hist_rad, hist_inc;
  int history_enable(int64_t cap) override {
    if (cap < 0) return IMC_ERR_ARG;
    hist_cap = cap; hist_dropped = 0; hist_temp.clear(); hist_mat.clear(); hist_rad.clear(); hist_inc.clear();
    return IMC_OK;
  }
  int history_count(int64_t* stored, int64_t* dropped) override {
    if (stored) *stored = (int64_t)hist_temp.size();
    if (dropped) *dropped = hist_dropped;
    return IMC_OK;
  }
  int history_get(int f, int64_t first, int64_t count, void* dst, int64_t bytes) override {
    if (first < 0 || count < 0 || first + count > (int64_t)hist_temp.size()) return IMC_ERR_ARG;
    if (f == IMC_FIELD_TEMP) {
      if (bytes != count * (int64_t)nc * 8) return IMC_ERR_ARG;
      for (int64_t s = 0; s < count; ++s) memcpy(static_cast<double*>(dst) + s * nc, hist_temp[first + s].data(), nc * 8);
      return IMC_OK;
    }
    const std::vector<std::vector<N>>* src = f == IMC_FIELD_MATENERGYDENS ? &hist_mat : f == IMC_FIELD_RADENERGYDENS ? &hist_rad
                                             : f == IMC_FIELD_NRG_INC ? &hist_inc : nullptr;
    if (!src || bytes != count * (int64_t)nc * P::bytes) return IMC_ERR_ARG;
    for (int64_t s = 0; s < count; ++s)
      for (size_t i = 0; i < nc; ++i) pack_native((*src)[first + s][i].d(), dst, (size_t)s * nc + i, P::bytes);
    return IMC_OK;
  }
  int history_clear() override { hist_dropped = 0; hist_temp.clear(); hist_mat.clear(); hist_rad.clear(); hist_inc.clear(); return IMC_OK; }
  int64_t num_particles() override { return (int64_t)particles.size(); }
  int get_particles(double* slots, uint64_t* ids_out, int64_t cap) override {
    if (cap < (int64_t)particles.size()) return IMC_ERR_ARG;
    for (size_t i = 0; i < particles.size(); ++i) {
      for (int k = 0; k < nslots; ++k) slots[i * nslots + k] = particles[i][k].d();
      if (ids_out) ids_out[i] = ids[i];
    }
    return IMC_OK;
  }
  int set_particles(const double* slots, const uint64_t* ids_in, int64_t n) override {
    particles.resize((size_t)n); ids.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      for (int k = 0; k < nslots; ++k) particles[i][k] = N::from_d(slots[i * nslots + k]);
      ids[i] = ids_in ? ids_in[i] : (uint64_t)i;
    }
    return IMC_OK;
  }
  int set_transport_tape(const double* u, int nu, const double* e, int ne, int64_t slots) override {
    tt_uni.assign(u, u + (size_t)nu * slots); tt_exp.assign(e, e + (size_t)ne * slots); tt_nuni = nu; tt_nexp = ne; tt_slots = slots;
    return IMC_OK;
  }
  int set_source_tape(const double* u, int nu, int64_t slots) override {
    st_uni.assign(u, u + (size_t)nu * slots); st_nuni = nu; st_slots = slots;
    return IMC_OK;
  }
  // ---- restart point (include/imc.h imc_checkpoint): what `deepcopy(mesh), deepcopy(particles), deepcopy(simvars)` would
  // give a Julia host — a copy of this whole object
  Oracle* ckpt = nullptr;
  ~Oracle() override { delete ckpt; }
  int checkpoint(int op) override {
    Oracle* held = ckpt;
    ckpt = nullptr;                                  // the copy must not own (or copy) a restart point itself
    if (op == IMC_CKPT_SAVE) { delete held; held = new Oracle(*this); }
    else if (op == IMC_CKPT_RESTORE) {
      if (!held) { err = "checkpoint: nothing saved"; return IMC_ERR_STATE; }
      *this = *held;
    } else if (op == IMC_CKPT_DROP) { delete held; held = nullptr; }
    else { ckpt = held; err = "checkpoint: unknown op"; return IMC_ERR_ARG; }
    ckpt = held;
    return IMC_OK;
  }
  // ---- Sourcing.sample_planck (imc_sourcing.jl:372-399): Fleck-Cummings series method.  Never called by the reference's
  // step (call sites commented out at :171, :186, :209, :232, :342, :362).  Sample i draws from Philox stream
  // (seed, id = i, step, STREAM_PLANCK) or from slot i of the source tape.
  int sample_planck(int64_t n_samples, int64_t step, double* out) override {
    if (n_samples < 0 || (n_samples > 0 && !out)) { err = "sample_planck: bad arguments"; return IMC_ERR_ARG; }
    const bool tape = cfg.rng_mode == IMC_RNG_TAPE;
    if (tape && n_samples > st_slots) { err = "source tape has fewer slots than samples"; return IMC_ERR_TAPE; }
    bool over = false;
    for (int64_t i = 0; i < n_samples; ++i) {
      PhiloxDraw<P> ph; TapeDraw<P> tp;
      if (tape) tp.init(st_uni.data(), st_nuni, nullptr, 0, (size_t)st_slots, (size_t)i);
      else ph.init((uint64_t)cfg.seed, (uint64_t)i, (uint32_t)step, STREAM_PLANCK);
      auto rnd = [&]() { return tape ? tp.uniform() : ph.uniform(); };
      N n = N::from_d(1.0);                                   // :382
      N rn1 = rnd();                                          // :383
      N nsum = N::from_d(1.0);                                // :384
      const double pi = 3.141592653589793;
      const double pi_4 = (pi * pi) * (pi * pi);              // pi^4 -> Float64 power by squaring
      double freq = std::nan("");
      for (int terms = 0; terms < 100000; ++terms) {          // `while true` in the reference: see include/imc.h for the cap
        if (rn1.d() <= 90.0 * nsum.d() / pi_4) {              // :388
          rn1 = rnd();                                        // :389
          N rn2 = rnd();                                      // :390
          N rn3 = rnd();                                      // :391
          N rn4 = rnd();                                      // :392
          N l = M::template log<P>(rn1 * rn2 * rn3 * rn4);
          freq = N::from_d(-1.0 * l.d() / n.d()).d();         // :393
          break;
        }
        n = n + N::from_d(1.0);                               // :396
        double n4;                                            // n^4 in T (Float16: through Float32, rounded once)
        if constexpr (P::id == 0) { float f = (float)n.d(); n4 = N::from_d((double)(f * f * f * f)).d(); }
        else { N sq = n * n; n4 = (sq * sq).d(); }
        nsum = nsum + N::from_d(1.0 / n4);                    // :397
      }
      out[i] = freq;
      if (tape && tp.exhausted()) over = true;
    }
    if (over) { err = "source tape exhausted"; return IMC_ERR_TAPE; }
    return IMC_OK;
  }
  int get_outcomes(int32_t* ev, int32_t* nseg, int64_t cap) override {
    if (cap < (int64_t)out_event.size()) return IMC_ERR_ARG;
    for (size_t i = 0; i < out_event.size(); ++i) { if (ev) ev[i] = out_event[i]; if (nseg) nseg[i] = out_nseg[i]; }
    return IMC_OK;
  }
};

}  // namespace imc_oracle
