// imc_kernels.cuh — the transport-step kernels, templated on the deck precision P (F16 / F32 / F64).
//
//   k_update          Update.update                 imc_update.jl:23-54          one thread per cell
//   k_src_energies    Sourcing.sourcing energies    imc_sourcing.jl:56-107       one thread per cell / edge
//   k_src_counts      particle counts per entry     imc_sourcing.jl:128-157,240-263
//   k_src_emit        particle creation             imc_sourcing.jl:159-236,264-366  one thread per new particle
//   k_track1d         Transport.MC                  imc_transport.jl:46-197      history-based, grid-stride
//   k_track2d         Transport.MC2D                imc_transport.jl:515-720
//   k_track1d_rw      Transport.MC_RW               imc_transport.jl:247-467
//   k_alive_count / k_compact   Clean.clean         imc_clean.jl:13-18           stable stream compaction
//   k_census_tally    Tally.tally census loop       imc_tally.jl:84-113
//   k_tally_finish    Tally.tally per-cell update   imc_tally.jl:29-32,44-75
//
// Arithmetic is Num<P>: one rounding per Julia operation, Float64 where the reference leaks into
// Float64.  Compiled with -fmad=false.  Particle state is structure-of-arrays.
#pragma once
#include <type_traits>
#include "imc_device.cuh"
#include "imc_fastdiv.cuh"
#include "imc_warp_reduce.cuh"
#include "imc_warp_runs.cuh"

namespace imc {

#ifndef IMC_TRACK_THREADS
#define IMC_TRACK_THREADS 256
#endif
#ifndef IMC_EARLY_NEXT_CELL
#define IMC_EARLY_NEXT_CELL 1   // MC2D: request the next cell's constants as soon as the event is known (0: in the face branch)
#endif
#ifndef IMC_CELL_GATHER_CG
#define IMC_CELL_GATHER_CG 0    // MC2D: 1 = gather the per-cell constants through L2 only; 0 = cached in L1 (the table is laid out so that
                                // consecutive cells of a history mostly share a sector, MeshDev::csx)
#endif
#ifndef IMC_FASTDIV_DIR
#define IMC_FASTDIV_DIR 1       // MC2D: divide by the direction cosines with cached reciprocals (imc_fastdiv.cuh)
#endif
#ifndef IMC_TRACK_MIN_BLOCKS
#define IMC_TRACK_MIN_BLOCKS 4
#endif
constexpr int TRACK_THREADS = IMC_TRACK_THREADS;
constexpr int QUEUE_CHUNK = 32;   // particles per claim of the dynamic schedule's queue
// resident blocks per SM the history kernels are compiled for: 4 x 256 threads (64 registers) for Float16 / Float32; Float64
// histories hold twice the registers — 3 blocks (80 registers) spill less and measured 5 % faster (crookedpipe_f64)
#ifndef IMC_TRACK_MIN_BLOCKS_F64
#define IMC_TRACK_MIN_BLOCKS_F64 3
#endif
#ifndef IMC_UNROLL2_F64
#define IMC_UNROLL2_F64 0   // Float64 histories make one Philox block per segment (no parity to exploit) and the second copy of the
                            // segment body costs registers at the 80-register cap: one segment per loop trip (crookedpipe_f64: 23.5 vs 29.3 ms)
#endif
template <class P> constexpr int track_min_blocks() { return P::id == 2 ? IMC_TRACK_MIN_BLOCKS_F64 : IMC_TRACK_MIN_BLOCKS; }

// reduce-buffer scalar slots that follow [energydep Nc*Ns | radenergydens Nc]
enum { RB_LOST = 0, RB_SEG, RB_HIST, RB_CENSUS, RB_ABSORBED, RB_ESCAPED, RB_RW, RB_ERRORS, RB_NSCALARS };

template <class P> struct CellProp1 { typename P::store_t w, dx, sig_col, neg_saf; };   // w = dx*ds
template <class P> struct alignas(2 * sizeof(typename P::store_t)) CellProp2 { typename P::store_t sig_col, neg_saf; };
// per-axis-index constants {w = d*ds, q}: q = 1/d in the table the ATOMIC / FIXED tallies read (deposit = E (1/dx)(1/dy) ..),
// q = d in the table the EXACT tallies read (deposit = (E/dx)/dy as the reference writes it); one 2-element load
template <class P> struct alignas(2 * sizeof(typename P::store_t)) AxisProp { typename P::store_t w, q; };

// The per-cell constants are gathered by cell index, once per cell entered, as one vector load.
template <class P>
__device__ __forceinline__ CellProp2<P> load_cell2(const CellProp2<P>* tab, long long c) {
  CellProp2<P> r;   // one 4 / 8 / 16-byte load
#if IMC_CELL_GATHER_CG
  if constexpr (P::id == 0) { const unsigned v = __ldcg(reinterpret_cast<const unsigned*>(tab) + c); r.sig_col = (uint16_t)(v & 0xffffu); r.neg_saf = (uint16_t)(v >> 16); }
  else if constexpr (P::id == 1) { const float2 v = __ldcg(reinterpret_cast<const float2*>(tab) + c); r.sig_col = v.x; r.neg_saf = v.y; }
  else { const double2 v = __ldcg(reinterpret_cast<const double2*>(tab) + c); r.sig_col = v.x; r.neg_saf = v.y; }
#else
  if constexpr (P::id == 0) { const unsigned v = reinterpret_cast<const unsigned*>(tab)[c]; r.sig_col = (uint16_t)(v & 0xffffu); r.neg_saf = (uint16_t)(v >> 16); }
  else if constexpr (P::id == 1) { const float2 v = reinterpret_cast<const float2*>(tab)[c]; r.sig_col = v.x; r.neg_saf = v.y; }
  else { const double2 v = reinterpret_cast<const double2*>(tab)[c]; r.sig_col = v.x; r.neg_saf = v.y; }
#endif
  return r;
}

// a / b for a divisor whose refined reciprocal r is cached (imc_fastdiv.cuh): Float16 / Float32 only — Float16 divides in
// Float32 and rounds, as Julia does; Float64 keeps the plain division
template <class P> __device__ __forceinline__ typename P::comp_t recip_of(Num<P> b) {
#if IMC_FASTDIV_DIR
  if constexpr (P::id == 2) return 0.0; else return FastDivisor::recip(b.v);
#else
  return 0;
#endif
}
template <class P> __device__ __forceinline__ Num<P> div_cached(Num<P> a, Num<P> b, typename P::comp_t r) {
  if constexpr (P::id == 2) return a / b;
  else { FastDivisor d; d.b = b.v; d.r = r; return Num<P>(P::rnd(d.divide(a.v))); }
}
// The two face distances of a 2-D segment, ax / vx and ay / vy, by the cached reciprocals with ONE shared range test
// (a slow-path region per division cost a BSSY / BRA / BSYNC triple and its call set-up each): both quotients are
// recomputed by the plain division when either numerator leaves [2^-40, 2^40] or the direction was marked slow
// (rvx == 0: load2d / the collision and reflection branches zero BOTH reciprocals when either cosine is out of range).
template <class P>
__device__ __forceinline__ void div_dir2(Num<P> ax, Num<P> vx, typename P::comp_t rvx, Num<P> ay, Num<P> vy, typename P::comp_t rvy, Num<P>& qx, Num<P>& qy) {
#if IMC_FASTDIV_DIR
  if constexpr (P::id != 2) {
    float q0 = __fmul_rn(ax.v, rvx), q1 = __fmul_rn(ay.v, rvy);
    q0 = __fmaf_rn(rvx, __fmaf_rn(-vx.v, q0, ax.v), q0);
    q1 = __fmaf_rn(rvy, __fmaf_rn(-vy.v, q1, ay.v), q1);
    const unsigned tx = (__float_as_uint(ax.v) & 0x7fffffffu) - 0x2b800000u, ty = (__float_as_uint(ay.v) & 0x7fffffffu) - 0x2b800000u;
    if (__builtin_expect(!(max(tx, ty) < (0x53ffffffu - 0x2b800000u) && rvx != 0.0f), 0)) { q0 = fastdiv_slow(ax.v, vx.v); q1 = fastdiv_slow(ay.v, vy.v); }
    qx = Num<P>(P::rnd(q0)); qy = Num<P>(P::rnd(q1));
    return;
  }
#endif
  qx = ax / vx; qy = ay / vy;
}
// both-or-none: the reciprocals of a direction (vx, vy) for div_dir2
template <class P>
__device__ __forceinline__ void recip_dir(Num<P> vx, Num<P> vy, typename P::comp_t& rvx, typename P::comp_t& rvy) {
  rvx = recip_of(vx); rvy = recip_of(vy);
  if constexpr (P::id != 2) { if (rvx == 0.0f || rvy == 0.0f) rvx = rvy = 0.0f; }
}
// division by a direction cosine (reciprocal cached in the history registers) — IMC_FASTDIV_DIR=0: plain division
template <class P> __device__ __forceinline__ Num<P> div_dir(Num<P> a, Num<P> b, typename P::comp_t r) {
#if IMC_FASTDIV_DIR
  return div_cached(a, b, r);
#else
  return a / b;
#endif
}

template <class P>
struct MeshDev {
  using S = typename P::store_t;
  using Cc = typename P::comp_t;
  int geom, nx, ny, ns;
  long long nc;
  S *dx, *dy, *wx, *wy;
  S *sa, *ss, *fleck, *beta, *bee, *sa_c, *sa_p, *ss_c, *ss_p, *sigma_static, *radsource;
  double* temp;
  S *matenergydens, *radenergydens, *nrg_inc, *energydep, *emittedenergy;
  S* tsurf[4];  // bottom, top, left, right (1-D: left = [2][0], right = [3][0])
  CellProp1<P>* cp1;
  CellProp2<P>* cp2;
  // Index maps of the two tables the 2-D history kernels touch once per segment by cell: index = xi*sx + yi*sy.
  // The fields themselves are column-major [xindex, yindex] (x fastest, like Julia); these internal tables are laid
  // out with the axis that particles cross MORE OFTEN (the one with the smaller mean cell width) fastest, so that
  // consecutive cells of a history mostly share a 32-byte sector: the gather of the next cell's constants then hits
  // L1 and consecutive deposits go to one L2 sector.  csx/csy: CellProp2 table (fixed at set_mesh); tsx/tsy: the
  // deposit accumulators of this launch (linear for EXACT records and for shared-memory accumulators).
  int csx, csy, tsx, tsy;
  AxisProp<P>*ax_inv, *ax_d;   // [nx + ny]: x entries, then y entries
  int ds_is_one, c_is_one;  // x / 1 == x exactly: the divisions by distancescale / phys_c can be skipped
  int n_tdiv; Cc tdiv[2], tdiv_r[2];   // tdiv_r: cached reciprocals (0 = take the plain division)   // the divisors of `(dist / ds) / c` that are not 1, in that order (a counted loop: the compiler
                            // turns `if (!is_one) x = x / d` into a division plus a select)
  Cc scales[IMC_MAX_SCALES];
  double scales_d[IMC_MAX_SCALES];
  Cc ds, c, a, alpha;
  int bc[4];
};

template <class P>
struct Parts {
  using S = typename P::store_t;
  S *t, *x, *y, *mu, *E, *E0;
  int *cx, *cy;             // 0-based cell indices (cy unused in 1-D)
  int* origin;              // slot 1 of the 1-D layout (never read by the physics)
  unsigned char* ks;        // energy-scale plane
  unsigned long long* id;   // Philox counter
};

struct RngArgs {
  int tape;
  unsigned long long seed;
  unsigned int step;
  unsigned int rk[20];   // Philox round keys of `seed` (Philox::round_keys)
  const double *uni, *ex;
  int n_uni, n_exp;
  long long stride;
};

// runtime-selected draw source (Philox or replay tape) for sourcing and MC_RW: sequential word consumption
template <class P>
struct Draw {
  int tape;
  PhiloxDraw<P> ph;
  TapeDraw<P> tp;
  __device__ __forceinline__ void init(const RngArgs& r, unsigned long long id, unsigned int stream, long long slot) {
    tape = r.tape;
    if (tape) tp.init(r.uni, r.n_uni, r.ex, r.n_exp, (size_t)r.stride, (size_t)slot);
    else ph.init(r.seed, id, r.step, stream);
  }
  __device__ __forceinline__ Num<P> uniform() { return tape ? tp.uniform() : ph.uniform(); }
  __device__ __forceinline__ Num<P> randexp() { return tape ? tp.randexp() : ph.randexp(); }
  __device__ __forceinline__ double randexp64() { return tape ? tp.randexp64() : ph.randexp64(); }
  __device__ __forceinline__ bool over() const { return tape && tp.exhausted(); }
};
// draw source of the MC / MC2D history loops, chosen at compile time: Philox words reserved per segment
// (SegDraw) or the replay tape.  Only the needed state lives in registers.
// `seg` is the 0-based segment index of the history (the caller's counter).
template <class P, bool TAPE, bool LEAN = false> struct HistDraw;
// PAR: parity of the segment index when the caller's loop is unrolled by two segments (-1: read it from `seg`)
template <class P>
struct HistDraw<P, false, false> {
  SegDraw<P> sg;
  __device__ __forceinline__ void init(const RngArgs&, unsigned long long id, long long) { sg.init(id); }
  __device__ __forceinline__ void resume(unsigned long long id, unsigned next, unsigned extra) { sg.resume(id, next, extra); }
  template <int PAR = -1> __device__ __forceinline__ void next_segment(const RngArgs& r, unsigned) { sg.next_segment_rk(r.rk, r.step); }
  template <int PAR = -1> __device__ __forceinline__ Num<P> uniform(const RngArgs& r, unsigned) { return sg.uniform(r.seed, r.step); }
  template <int PAR = -1> __device__ __forceinline__ Num<P> randexp(unsigned) { return sg.randexp(); }
  __device__ __forceinline__ bool over() const { return false; }
};
template <class P>
struct HistDraw<P, false, true> {   // MC2D history kernels
  SegDrawLean<P> sg;
  __device__ __forceinline__ void init(const RngArgs&, unsigned long long id, long long) { sg.init(id); }
  template <int PAR = -1> __device__ __forceinline__ void next_segment(const RngArgs& r, unsigned seg) { sg.template next_segment_rk<PAR>(r.rk, r.step, seg); }
  template <int PAR = -1> __device__ __forceinline__ Num<P> uniform(const RngArgs&, unsigned seg) { return sg.template uniform<PAR>(seg); }
  template <int PAR = -1> __device__ __forceinline__ Num<P> randexp(unsigned seg) { return sg.template randexp<PAR>(seg); }
  __device__ __forceinline__ bool over() const { return false; }
};
template <class P, bool LEAN>
struct HistDraw<P, true, LEAN> {
  TapeDraw<P> tp;
  __device__ __forceinline__ void init(const RngArgs& r, unsigned long long, long long slot) { tp.init(r.uni, r.n_uni, r.ex, r.n_exp, (size_t)r.stride, (size_t)slot); }
  template <int PAR = -1> __device__ __forceinline__ void next_segment(const RngArgs&, unsigned) {}
  template <int PAR = -1> __device__ __forceinline__ Num<P> uniform(const RngArgs&, unsigned) { return tp.uniform(); }
  template <int PAR = -1> __device__ __forceinline__ Num<P> randexp(unsigned) { return tp.randexp(); }
  __device__ __forceinline__ bool over() const { return tp.exhausted(); }
};

// ======================================================================================
// Update.update
// ======================================================================================
template <class P>
__global__ void k_update(MeshDev<P> m, typename P::comp_t dt_, int linearized, int marshak, int temp_wide) {
  using N = Num<P>;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.nc) return;
  N dt(dt_), a(m.a), one = N::from_d(1.0);
  double t = m.temp[i];
  N fourA = N::from_i(4) * a;
  N val; double w = 0;
  if (temp_wide) { w = fourA.d() * ((t * t) * t); val = N::from_d(w); }
  else { N tt = N::from_d(t); val = fourA * ((tt * tt) * tt); }
  N bee = N::load(m.bee, i), beta;
  if (linearized) { bee = val; bee.store(m.bee, i); beta = one; }                    // :24-25
  else beta = temp_wide ? N::from_d(w / bee.d()) : val / bee;                       // :27
  beta.store(m.beta, i);
  N sa_c = N::load(m.sa_c, i), sa_p = N::load(m.sa_p, i), ss_c = N::load(m.ss_c, i), ss_p = N::load(m.ss_p, i);
  N sa, ss;
  if (temp_wide) {
    sa = N::from_d(sa_c.d() * MathDet::pow64(t, sa_p.d()));
    if (m.geom == 1 && marshak) sa = N::from_d(((sa_c.d() / t) / t) / t);          // Q18
    ss = N::from_d(ss_c.d() * MathDet::pow64(t, ss_p.d()));
  } else {
    N tt = N::from_d(t);
    sa = sa_c * MathDet::pow<P>(tt, sa_p);
    if (m.geom == 1 && marshak) sa = ((sa_c / tt) / tt) / tt;
    ss = ss_c * MathDet::pow<P>(tt, ss_p);
  }
  sa.store(m.sa, i); ss.store(m.ss, i);
  double vals[6] = {(double)m.ds, (double)m.alpha, beta.d(), (double)m.c, dt.d(), sa.d()};
  double sc1 = 1.0;
  N prod; int idx;
  sorter_dev<P, 6>(vals, &sc1, 1, &prod, &idx);
  N f = N::from_d(1.0 / (1.0 + prod.d()));                                         // :39 / :52
  f.store(m.fleck, i);
  // per-cell quantities the tracking loop recomputes every segment in the reference, formed once
  // here with the same operations and roundings (imc_transport.jl:87, :95)
  N sig_col = sa * (one - f) + ss;
  N neg_saf = (-sa) * f;
  if (m.geom == 1) {
    CellProp1<P> c;
    N dx = N::load(m.dx, i);
    c.w = P::pack((dx * N(m.ds)).v); c.dx = P::pack(dx.v); c.sig_col = P::pack(sig_col.v); c.neg_saf = P::pack(neg_saf.v);
    m.cp1[i] = c;
  } else {
    CellProp2<P> c;
    c.sig_col = P::pack(sig_col.v); c.neg_saf = P::pack(neg_saf.v);
    m.cp2[m.geom == 2 ? (i % m.nx) * (long long)m.csx + (i / m.nx) * (long long)m.csy : i] = c;
  }
}

// wx = dx*ds, wy = dy*ds (static)
template <class P>
__global__ void k_widths(MeshDev<P> m) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  using N = Num<P>;
  if (i < m.nx) {
    N d = N::load(m.dx, i), w = d * N(m.ds);
    w.store(m.wx, i);
    AxisProp<P> a; a.w = P::pack(w.v); a.q = P::pack((N::from_d(1.0) / d).v); m.ax_inv[i] = a;
    a.q = P::pack(d.v); m.ax_d[i] = a;
  }
  if (m.geom == 2 && i < m.ny) {
    N d = N::load(m.dy, i), w = d * N(m.ds);
    w.store(m.wy, i);
    AxisProp<P> a; a.w = P::pack(w.v); a.q = P::pack((N::from_d(1.0) / d).v); m.ax_inv[m.nx + i] = a;
    a.q = P::pack(d.v); m.ax_d[m.nx + i] = a;
  }
}

// ======================================================================================
// Sourcing
// ======================================================================================
// Entry layout.  1-D: [left, right, body(Nc), radsource(Nc)];  2-D: [bottom(Nx), top(Nx), left(Ny),
// right(Ny), body(Nc), radsource(Nc)] — the reference's emission order (imc_sourcing.jl:161-236, :265-366).
struct SrcLayout {
  int geom, nx, ny;
  long long nc;
  __host__ __device__ long long n_surf() const { return geom == 1 ? 2 : 2ll * nx + 2ll * ny; }
  __host__ __device__ long long body0() const { return n_surf(); }
  __host__ __device__ long long rad0() const { return n_surf() + nc; }
  __host__ __device__ long long total() const { return n_surf() + 2 * nc; }
};

template <class P>
struct SrcArrays {
  using S = typename P::store_t;
  S* e;               // energy product per entry (already multiplied by its scale)
  S* q;               // e / escale (for the totalenergy sums)
  signed char* ks;    // chosen scale plane per entry (-1: sorter failed)
  int* cnt;           // loop count per entry
  S* nrg;             // energy per particle of the entry
  S* q_em;            // emittedenergy ./ escale, [Nc x Ns], for the print at :75
};

struct SrcScalars {   // written by k_src_total, read by k_src_counts and by the host
  double totalenergy; // value of T
  double nsrc;        // n_source as a T value
  double sums[8];
  int bad;
};

template <class P>
__global__ void k_src_energies(MeshDev<P> m, SrcArrays<P> s, SrcLayout L, typename P::comp_t dt_) {
  using N = Num<P>;
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= L.n_surf() + L.nc) return;  // one thread per surface entry and per cell (body + radsource + emitted)
  const double dt = (double)dt_, a = (double)m.a, c = (double)m.c, ds = (double)m.ds;
  N prod; int idx;
  if (e < L.n_surf()) {
    if (L.geom == 1) {                                                            // :60-61
      double ts = Num<P>::load(m.tsurf[e == 0 ? 2 : 3], 0).d();
      const double vals[8] = {a, c, ts, ts, ts, ts, dt, 0.25};
      sorter_dev<P, 8>(vals, m.scales_d, m.ns, &prod, &idx);
    } else {                                                                      // :86-93
      int side; long long i = e;
      if (i < L.nx) side = 0; else if ((i -= L.nx) < L.nx) side = 1; else if ((i -= L.nx) < L.ny) side = 2; else { i -= L.ny; side = 3; }
      double ts = Num<P>::load(m.tsurf[side], i).d();
      double width = side < 2 ? Num<P>::load(m.dx, i).d() : Num<P>::load(m.dy, i).d();
      const double vals[9] = {a, c, ts, ts, ts, ts, width, dt, 0.25};
      sorter_dev<P, 9>(vals, m.scales_d, m.ns, &prod, &idx);
    }
    prod.store(s.e, e); s.ks[e] = (signed char)idx;
    (prod / (idx >= 0 ? N(m.scales[idx]) : N())).store(s.q, e);
    return;
  }
  long long i = e - L.n_surf();
  int xi = (int)(L.geom == 1 ? i : i % L.nx), yi = (int)(L.geom == 1 ? 0 : i / L.nx);
  double t = m.temp[i];
  double f = N::load(m.fleck, i).d(), sa = N::load(m.sa, i).d(), dx = N::load(m.dx, xi).d();
  double dy = L.geom == 2 ? N::load(m.dy, yi).d() : 0.0;
  double rs = N::load(m.radsource, i).d();
  {                                                                               // body :67 / :102
    if (L.geom == 2) { const double vals[12] = {f, sa, a, c, t, t, t, t, dx, dy, dt, ds}; sorter_dev<P, 12>(vals, m.scales_d, m.ns, &prod, &idx); }
    else { const double vals[11] = {f, sa, a, c, t, t, t, t, dx, dt, ds}; sorter_dev<P, 11>(vals, m.scales_d, m.ns, &prod, &idx); }
    long long eb = L.body0() + i;
    prod.store(s.e, eb); s.ks[eb] = (signed char)idx;
    (prod / (idx >= 0 ? N(m.scales[idx]) : N())).store(s.q, eb);
  }
  {                                                                               // radiation source :68 / :103
    if (L.geom == 2) { const double vals[4] = {rs, dx, dy, dt}; sorter_dev<P, 4>(vals, m.scales_d, m.ns, &prod, &idx); }
    else { const double vals[3] = {rs, dx, dt}; sorter_dev<P, 3>(vals, m.scales_d, m.ns, &prod, &idx); }
    long long er = L.rad0() + i;
    prod.store(s.e, er); s.ks[er] = (signed char)idx;
    (prod / (idx >= 0 ? N(m.scales[idx]) : N())).store(s.q, er);
  }
  {                                                                               // emitted energy density :69-70 / :104-105
    double vals[10] = {f, sa, a, c, t, t, t, t, dt, ds};
    sorter_dev<P, 10>(vals, m.scales_d, m.ns, &prod, &idx);
    for (int k = 0; k < m.ns; ++k) {
      N v = (k == idx) ? prod : N();
      v.store(m.emittedenergy, i + m.nc * k);
      (v / (idx >= 0 ? N(m.scales[idx]) : N())).store(s.q_em, i + m.nc * k);
    }
  }
}

// totalenergy and n_source (imc_sourcing.jl:121, :132-136); sums[] hold the jl_sum results:
// 1-D: [body, rad]; 2-D: [bottom, top, left, right, body, rad]
template <class P>
__global__ void k_src_total(SrcArrays<P> s, SrcLayout L, const typename P::comp_t* sums, SrcScalars* out,
                            long long n_input, long long n_census, long long n_max, typename P::comp_t cellmin, int wide_counts) {
  using N = Num<P>;
  N e_surface, body, rad;
  if (L.geom == 1) {
    e_surface = N::load(s.q, 0) + N::load(s.q, 1);                                // :64
    body = N(sums[0]); rad = N(sums[1]);
  } else {
    e_surface = ((N(sums[0]) + N(sums[1])) + N(sums[2])) + N(sums[3]);            // :95
    body = N(sums[4]); rad = N(sums[5]);
  }
  N total = (body + rad) + e_surface;                                             // :121
  out->totalenergy = total.d();
  auto toT = [&](long long v) { return wide_counts ? (double)(float)v : N::from_i(v).d(); };
  double nsrc = toT(n_input);
  if (n_input + n_census > n_max) {                                               // :134-136 (Q9)
    long long cand = n_max - n_census - (L.geom == 1 ? 1 : 2) - 1;
    double cand_t = toT(cand);
    nsrc = (double)cellmin > cand_t ? (double)cellmin : cand_t;
  }
  out->nsrc = nsrc;
  out->bad = 0;
}

template <class P>
__device__ __forceinline__ long long count_of(Num<P> e, Num<P> escale, double nsrc, double total, Num<P> cellmin,
                                               bool floor_cellmin, int wide_counts, int* bad) {
  double r;
  if (wide_counts) {  // Float16 decks with counts beyond Float16: Float32 arithmetic (Q10)
    using W = Num<F32>;
    W x = ((W((float)e.v) / W((float)escale.v)) * W::from_d(nsrc)) / W((float)total);
    x = jl_round(x);
    if (floor_cellmin) x = jl_max(x, W((float)cellmin.v));
    r = x.d();
  } else {
    Num<P> x = ((e / escale) * Num<P>::from_d(nsrc)) / Num<P>::from_d(total);
    x = jl_round(x);
    if (floor_cellmin) x = jl_max(x, cellmin);
    r = x.d();
  }
  if (!(r - r == 0.0) || r < 0 || r > 2147483647.0) { *bad = 1; return 0; }
  return (long long)r;
}
template <class P>
__device__ __forceinline__ Num<P> div_count(Num<P> e, long long n) {
  if constexpr (P::id == 0) { if (n > 65504) return Num<P>(P::rnd(e.v / (float)n)); }
  return e / Num<P>::from_i(n);
}

template <class P>
__global__ void k_src_counts(MeshDev<P> m, SrcArrays<P> s, SrcLayout L, SrcScalars* sc, typename P::comp_t cellmin_, int wide_counts) {
  using N = Num<P>;
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= L.total()) return;
  N cellmin(cellmin_);
  const double nsrc = sc->nsrc, total = sc->totalenergy;
  int bad = 0;
  N en = N::load(s.e, e);
  int ks = s.ks[e];
  N escale = ks >= 0 ? N(m.scales[ks]) : N();
  long long cnt = 0, loops = 0;
  if (e < L.n_surf()) {                                                            // :150-157 (1-D, no floor) / :240-263 (2-D)
    if (en > N()) cnt = count_of(en, escale, nsrc, total, cellmin, L.geom == 2, wide_counts, &bad);
    loops = cnt;
  } else if (e < L.rad0()) {                                                       // :138-140
    cnt = count_of(en, escale, nsrc, total, cellmin, true, wide_counts, &bad);
    loops = cnt;
  } else {                                                                         // :142-146
    if (en > N()) cnt = count_of(en, escale, nsrc, total, cellmin, true, wide_counts, &bad);
    loops = cnt;
    if (L.geom == 2 && cnt > 0) {                                                  // Q6: the 2-D loop runs 1:n_body
      long long eb = e - L.nc;
      N enb = N::load(s.e, eb);
      int ksb = s.ks[eb];
      loops = count_of(enb, ksb >= 0 ? N(m.scales[ksb]) : N(), nsrc, total, cellmin, true, wide_counts, &bad);
    }
  }
  if (ks < 0) bad = 1;
  s.cnt[e] = (int)loops;
  (cnt > 0 ? div_count(en, cnt) : N()).store(s.nrg, e);
  if (bad) atomicOr(&sc->bad, 1);
}

// One new particle: entry e of the source list (surfaces, body cells, radiation-source cells), global ordinal j in the
// reference's emission order (particle id and Philox counter), slot o of the particle list.  Cell indices are 32-bit.
template <class P>
__device__ __forceinline__ void emit_one(const MeshDev<P>& m, const Parts<P>& p, const SrcArrays<P>& s, const SrcLayout& L, long long e,
                                         long long j, long long o, typename P::comp_t dt_, const RngArgs& rng, unsigned long long* over_flag) {
  using N = Num<P>;
  N dt(dt_), ds(m.ds), one = N::from_d(1.0), two = N::from_i(2);
  const double PI = 3.141592653589793;
  unsigned long long id = ((unsigned long long)rng.step << 40) | (unsigned long long)j;
  Draw<P> d; d.init(rng, id, STREAM_SOURCE, j);
  N t, x, y, mu, nrg = N::load(s.nrg, e);
  int cx = 0, cy = 0;
  const int nsurf = (int)L.n_surf(), nc = (int)L.nc;
  if (L.geom == 1) {
    if (e < 2) {                                                                   // :161-189
      bool left = e == 0;
      cx = left ? 0 : nc - 1;
      x = N::from_d(((left ? 0.01 : 0.99) * N::load(m.dx, cx).d()) * ds.d());
      mu = MathDet::sqrt<P>(d.uniform());
      if (!left) mu = -mu;
      while (mu == N()) { mu = MathDet::sqrt<P>(d.uniform()); if (!left) mu = -mu; }
      t = dt * d.uniform();
    } else {                                                                       // :193-236
      const int c = (int)(e - 2);
      cx = c < nc ? c : c - nc;
      x = (N::load(m.dx, cx) * d.uniform()) * ds;
      mu = one - two * d.uniform();
      while (mu == N()) mu = one - two * d.uniform();
      t = dt * d.uniform();
    }
  } else {
    if (e < nsurf) {                                                               // :265-323
      int side; int i = (int)e;
      if (i < L.nx) side = 0; else if ((i -= L.nx) < L.nx) side = 1; else if ((i -= L.nx) < L.ny) side = 2; else { i -= L.ny; side = 3; }
      t = dt * d.uniform();
      if (side == 0) {
        cx = i; cy = 0;
        x = (N::load(m.dx, i) * d.uniform()) * ds;
        y = N::from_d((0.001 * N::load(m.dy, 0).d()) * ds.d());
        mu = N::from_d(PI) * d.uniform();
      } else if (side == 1) {
        cx = i; cy = L.ny - 1;
        x = (N::load(m.dx, i) * d.uniform()) * ds;
        y = N::from_d((0.999 * N::load(m.dy, L.ny - 1).d()) * ds.d());
        mu = N::from_d((-PI) * d.uniform().d());
      } else {
        cx = side == 2 ? 0 : L.nx - 1; cy = i;
        x = N::from_d(((side == 2 ? 0.001 : 0.999) * N::load(m.dx, cx).d()) * ds.d());
        N dq7 = i < L.nx ? N::load(m.dx, i) : N::load(m.dy, i);                    // Q7: mesh.dx[j]
        y = (dq7 * d.uniform()) * ds;
        double u = d.uniform().d();
        mu = N::from_d(PI * (side == 2 ? 0.5 - u : 0.5 + u));
      }
    } else {                                                                       // :326-366
      int c = (int)(e - nsurf);
      if (c >= nc) c -= nc;
      cy = c / L.nx; cx = c - cy * L.nx;
      x = (N::load(m.dx, cx) * d.uniform()) * ds;
      y = (N::load(m.dy, cy) * d.uniform()) * ds;
      mu = N::from_d((2.0 * PI) * d.uniform().d());
      t = dt * d.uniform();
    }
  }
  t.store(p.t, o); x.store(p.x, o); mu.store(p.mu, o); nrg.store(p.E, o); nrg.store(p.E0, o);
  p.cx[o] = cx; p.ks[o] = (unsigned char)(s.ks[e] < 0 ? 0 : s.ks[e]); p.id[o] = id;
  if (L.geom == 2) { y.store(p.y, o); p.cy[o] = cy; } else p.origin[o] = cx;
  if (d.over()) atomicAdd(over_flag, 1ull);
}

// thread per new particle: the entry (source cell or surface element) of particle ordinal j is the last e with
// offs[e] <= j.  A full binary search in the scan of the counts is 25 dependent loads over 268 MB (4096^2 mesh: 64 % of
// the kernel's instructions, 79 % of its stall samples), so it is done once per BLOCK by a small pre-kernel
// (k_src_block_entries: entry of each block's first particle); inside the block the ordinals are consecutive, so every
// thread searches only between its block's entry and the next block's — a window of a few hundred entries at most
// (one entry when a hot surface element emits the whole block), resident in L1.  Robust for any count distribution.
// Measured alternatives of round 1, all slower or equal: thread per entry (16x slower on Marshak), one search per warp +
// shuffle window, one search per block inside the kernel.
constexpr int EMIT_THREADS = 256;
static __global__ void k_src_block_entries(const long long* __restrict__ offs, long long n_entries, long long n_local, int rank, int world,
                                           long long n_blocks, long long* __restrict__ block_entry) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_blocks) return;
  if (b == n_blocks) { block_entry[b] = n_entries - 1; return; }
  const long long j = rank + (b * EMIT_THREADS) * (long long)world;
  long long lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const long long mid = (lo + hi + 1) >> 1;
    if (offs[mid] <= j) lo = mid; else hi = mid - 1;
  }
  block_entry[b] = lo;
}
template <class P>
__global__ void __launch_bounds__(EMIT_THREADS) k_src_emit(MeshDev<P> m, Parts<P> p, SrcArrays<P> s, SrcLayout L, const long long* __restrict__ offs,
                           const long long* __restrict__ block_entry, long long base, long long n_local, int rank, int world, typename P::comp_t dt_,
                           RngArgs rng, unsigned long long* over_flag) {
  long long l = (long long)blockIdx.x * EMIT_THREADS + threadIdx.x;
  if (l >= n_local) return;
  long long j = rank + l * (long long)world;  // global ordinal in the reference's emission order
  // entry = last e with offs[e] <= j, between the entries of this block's and the next block's first particle
  long long lo = block_entry[blockIdx.x], hi = block_entry[blockIdx.x + 1];
  while (lo < hi) {
    long long mid = (lo + hi + 1) >> 1;
    if (offs[mid] <= j) lo = mid; else hi = mid - 1;
  }
  emit_one(m, p, s, L, lo, j, base + l, dt_, rng, over_flag);
}

// ======================================================================================
// Tallies
// ======================================================================================
template <class P> struct AccType { using type = float; };
template <> struct AccType<F64> { using type = double; };

struct TallyArgs {
  int mode;          // IMC_TALLY_ATOMIC or IMC_TALLY_FIXED (EXACT uses the record path)
  int use_smem;      // block-private accumulators in shared memory (nacc of them)
  int copies;        // shared-memory accumulator sets per block (a power of two): warp w deposits into set w % copies, so that
                     // the compare-and-swap loops of different warps do not collide
  int nacc;          // Nc * Ns
  double* g_acc;     // reduce buffer viewed as Float64 (ATOMIC)
  long long* g_fx;   // reduce buffer viewed as int64 (FIXED)
  double fx_mul;     // 2^S for densities
  double fx_mul_lost;
  long long sc0;     // index of the first scalar slot in the reduce buffer
  // EXACT mode (reference summation order): pass 1 counts the deposits of every particle, pass 2 writes one
  // record per deposit at rec_off[particle] + segment, i.e. in the order the reference's loops produce them
  int pass;                 // 0 = normal, 1 = count only (no side effects), 2 = write records
  int* rec_cnt;             // [n] deposits per particle (pass 1)
  const long long* rec_off; // [n] exclusive scan of rec_cnt
  unsigned* rec_key;        // [R] cell + Nc*k; bit 31 = the value is a Float64 deposit (MC_RW, Q3)
  double* rec_val;          // [R]
  double* lost_val;         // [n] energy lost through a VACUUM boundary (NaN = none)
};

// Tally kind of a tracking kernel.  TK_RUNTIME reads TallyArgs (EXACT passes, replay tape, MC_RW, event schedule,
// census tally); the four other kinds hard-wire the accumulator representation and its location, so the hot
// Philox history kernels carry no mode tests in the segment loop.
enum { TK_RUNTIME = -1, TK_ATOMIC_G = 0, TK_ATOMIC_S = 1, TK_FIXED_G = 2, TK_FIXED_S = 3 };
template <int TK>
struct TKind {
  static __device__ __forceinline__ bool exact(const TallyArgs& a) { if constexpr (TK == TK_RUNTIME) return a.mode == IMC_TALLY_EXACT; else return false; }
  static __device__ __forceinline__ bool fixed(const TallyArgs& a) { if constexpr (TK == TK_RUNTIME) return a.mode == IMC_TALLY_FIXED; else return TK == TK_FIXED_G || TK == TK_FIXED_S; }
  static __device__ __forceinline__ bool smem(const TallyArgs& a) { if constexpr (TK == TK_RUNTIME) return a.use_smem != 0; else return TK == TK_ATOMIC_S || TK == TK_FIXED_S; }
};

// 64-bit shared-memory accumulator += 64-bit value by native 32-bit atomics (low word, carry, high word): a 64-bit shared
// atomicAdd is a compare-and-swap loop on sm_100.  Two's-complement, so negative addends work; readers synchronise first.
__device__ __forceinline__ void smem_add64(unsigned long long* acc, unsigned long long q) {
  unsigned* w = reinterpret_cast<unsigned*>(acc);
  const unsigned lo = (unsigned)q, hi = (unsigned)(q >> 32);
  const unsigned old = atomicAdd(w, lo);
  const unsigned carry = old > 0xffffffffu - lo ? 1u : 0u;
  if (hi + carry) atomicAdd(w + 1, hi + carry);
}

template <class P, int TK = TK_RUNTIME>
struct Tally {
  using A = typename AccType<P>::type;
  using K = TKind<TK>;
  const TallyArgs& a;
  A* s_acc; unsigned long long* s_fx;
  A* s_acc0; unsigned long long* s_fx0;   // set 0 (zero / flush walk all sets from here)
  __device__ __forceinline__ Tally(const TallyArgs& a_, unsigned char* smem) : a(a_) {
    s_acc0 = reinterpret_cast<A*>(smem); s_fx0 = reinterpret_cast<unsigned long long*>(smem);
    const int set = K::smem(a_) ? (int)((threadIdx.x >> 5) & (unsigned)(a_.copies - 1)) : 0;
    s_acc = s_acc0 + (size_t)set * a_.nacc; s_fx = s_fx0 + (size_t)set * a_.nacc;
  }
  __device__ __forceinline__ void zero() {
    if (!K::smem(a) || K::exact(a)) return;
    const int n = a.nacc * a.copies;
    if (K::fixed(a)) { for (int i = threadIdx.x; i < n; i += blockDim.x) s_fx0[i] = 0ull; }
    else { for (int i = threadIdx.x; i < n; i += blockDim.x) s_acc0[i] = (A)0; }
    __syncthreads();
  }
  __device__ __forceinline__ void add(long long idx, Num<P> v, long long rec = 0, bool wide = false, double wide_v = 0.0) {
    if (K::exact(a)) {
      if (a.pass == 2) { a.rec_key[rec] = (unsigned)idx | (wide ? 0x80000000u : 0u); a.rec_val[rec] = wide ? wide_v : v.d(); }
      return;
    }
    if (K::fixed(a)) {
      long long q = __double2ll_rn(v.d() * a.fx_mul);
      if (K::smem(a)) smem_add64(&s_fx[idx], (unsigned long long)q);
      else atomicAdd(reinterpret_cast<unsigned long long*>(a.g_fx) + idx, (unsigned long long)q);
    } else {
      if (K::smem(a)) atomicAdd(&s_acc[idx], (A)v.v);
      else atomicAdd(a.g_acc + idx, v.d());
    }
  }
  // 1-D decks: the lanes of a warp mostly deposit into the same one or two cells, and a shared-memory Float32 / 64-bit
  // atomic add is a compare-and-swap loop (27 % of k_track1d's instructions on Su-Olson were its retries).  Runs of
  // adjacent active lanes with the same accumulator are therefore summed first with a segmented shuffle scan and the last
  // lane of each run issues the one atomic.  Float32 partial sums in ATOMIC mode (order-free to the tolerance, like the
  // per-lane atomics), exact integer sums in FIXED mode.  Only for the shared-memory kinds; everything else -> add().
  __device__ __forceinline__ void add_runs(int idx, Num<P> v, long long rec = 0) {
    if (!K::smem(a) || K::exact(a)) { add(idx, v, rec); return; }
    const unsigned act = __activemask();
    const int lane = threadIdx.x & 31;
    const int prev = __shfl_up_sync(act, idx, 1);
    const bool head = lane == 0 || !((act >> (lane - 1)) & 1u) || prev != idx;
    const unsigned heads = __ballot_sync(act, head);
    const int headlane = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
    const bool tail = lane == 31 || !((act >> (lane + 1)) & 1u) || ((heads >> (lane + 1)) & 1u);
    if (K::fixed(a)) {
      long long q = __double2ll_rn(v.d() * a.fx_mul);
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) { const long long t = __shfl_up_sync(act, q, dlt); if (lane - dlt >= headlane) q += t; }
      if (tail) smem_add64(&s_fx[idx], (unsigned long long)q);
    } else {
      A x = (A)v.v;
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) { const A t = __shfl_up_sync(act, x, dlt); if (lane - dlt >= headlane) x += t; }
      if (tail) atomicAdd(&s_acc[idx], x);
    }
  }
  __device__ __forceinline__ void flush() {
    if (!K::smem(a) || K::exact(a)) return;
    __syncthreads();
    if (K::fixed(a)) {
      for (int i = threadIdx.x; i < a.nacc; i += blockDim.x) {
        unsigned long long q = 0ull;
        for (int c = 0; c < a.copies; ++c) q += s_fx0[(size_t)c * a.nacc + i];
        if (q) atomicAdd(reinterpret_cast<unsigned long long*>(a.g_fx) + i, q);
      }
    } else {
      for (int i = threadIdx.x; i < a.nacc; i += blockDim.x) {
        double q = 0.0;
        for (int c = 0; c < a.copies; ++c) q += (double)s_acc0[(size_t)c * a.nacc + i];
        if (q != 0.0) atomicAdd(a.g_acc + i, q);
      }
    }
  }
};

// Event counters and lost energy live in shared memory, not in registers: they are touched once per history, and 18
// fewer live registers in the tracking loop buy a fourth resident block per SM.
//   * per THREAD: segments (64-bit) and the four outcome counts (32-bit; a thread ends fewer than 2^32 histories because
//     the engine refuses more than 2^32 particles per GPU).  Thread-private slots need no atomics, no votes and no
//     leader election: the end of a history costs three load-add-store triples.  (Round 1 kept one slot set per block
//     and updated it with warp-aggregated shared atomics: ~200 SASS instructions per write-back, 5 % of the kernel.)
//     Layout [counter][thread]: consecutive lanes hit consecutive words, conflict-free.
//   * per BLOCK: RB_LOST (a Float64 in ATOMIC mode, a fixed-point integer in FIXED mode), updated atomically (escapes are rare).
struct Counters {
  unsigned long long* s;     // block slots [RB_NSCALARS]
  unsigned long long* seg;   // [blockDim.x]
  unsigned* evn;             // [5][blockDim.x]: census, absorbed, escaped, random-walk kill; [4] = numeric errors
  double* park;              // [blockDim.x]: one value per thread that is needed only at the end of a history (Float64 MC2D: the angle)
  __device__ __forceinline__ void init(unsigned long long* slots) {
    s = slots;
    seg = slots + RB_NSCALARS + threadIdx.x;
    evn = reinterpret_cast<unsigned*>(slots + RB_NSCALARS + blockDim.x) + threadIdx.x;
    park = reinterpret_cast<double*>(evn - threadIdx.x + 5 * blockDim.x + (blockDim.x & 1u)) + threadIdx.x;   // 8-byte aligned
    if (threadIdx.x < RB_NSCALARS) s[threadIdx.x] = 0ull;
    *seg = 0ull;
#pragma unroll
    for (int k = 0; k < 5; ++k) evn[k * blockDim.x] = 0u;
    __syncthreads();
  }
  // end of one history: outcome ev (0 census, 1 absorbed, 2 escaped, 3 random-walk kill) after nseg segments
  __device__ __forceinline__ void finish(int ev, int nseg) {
    *seg += (unsigned long long)(unsigned)nseg;
    evn[ev * blockDim.x] += 1u;
  }
  __device__ __forceinline__ void error() { evn[4 * blockDim.x] += 1u; }
  template <int TK = TK_RUNTIME>
  __device__ __forceinline__ void lose_value(const TallyArgs& a, double e_over_scale) {
    if (TKind<TK>::fixed(a)) atomicAdd(&s[RB_LOST], (unsigned long long)__double2ll_rn(e_over_scale * a.fx_mul_lost));
    else atomicAdd(reinterpret_cast<double*>(&s[RB_LOST]), e_over_scale);
  }
  template <class P, int TK = TK_RUNTIME>
  __device__ __forceinline__ void lose(const TallyArgs& a, Num<P> e_over_scale, long long pi = 0, double e_raw = 0.0) {
    if (TKind<TK>::exact(a)) { if (a.pass == 2) a.lost_val[pi] = e_raw; return; }
    lose_value<TK>(a, e_over_scale.d());
  }
  template <int TK = TK_RUNTIME>
  __device__ __forceinline__ void commit(const TallyArgs& a) {
    __syncthreads();
    if (TKind<TK>::exact(a) && a.pass == 1) return;
    // per-thread slots -> one sum per warp -> the reduce buffer
    unsigned long long sg = *seg;
    unsigned c0 = evn[0], c1 = evn[blockDim.x], c2 = evn[2 * blockDim.x], c3 = evn[3 * blockDim.x], c4 = evn[4 * blockDim.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sg += __shfl_xor_sync(IMC_FULL_MASK, sg, o);
    c0 = __reduce_add_sync(IMC_FULL_MASK, c0); c1 = __reduce_add_sync(IMC_FULL_MASK, c1);
    c2 = __reduce_add_sync(IMC_FULL_MASK, c2); c3 = __reduce_add_sync(IMC_FULL_MASK, c3); c4 = __reduce_add_sync(IMC_FULL_MASK, c4);
    const int lane = threadIdx.x & 31;
    if (lane < RB_NSCALARS && lane != RB_LOST) {
      // per-warp sums < 2^32 * 32 fit 64 bits; counts are exact in Float64 below 2^53
      const unsigned long long hist = (unsigned long long)c0 + c1 + c2 + c3;
      const unsigned long long v = lane == RB_SEG ? sg : lane == RB_HIST ? hist : lane == RB_CENSUS ? (unsigned long long)c0
                                 : lane == RB_ABSORBED ? (unsigned long long)c1 + c3 : lane == RB_ESCAPED ? (unsigned long long)c2
                                 : lane == RB_RW ? (unsigned long long)c3 : (unsigned long long)c4;
      if (v) {
        if (TKind<TK>::fixed(a)) atomicAdd(reinterpret_cast<unsigned long long*>(a.g_fx) + a.sc0 + lane, v);
        else atomicAdd(a.g_acc + a.sc0 + lane, (double)v);
      }
    }
    const int k = threadIdx.x;
    if (k == RB_LOST && s[k] != 0ull) {
      if (TKind<TK>::fixed(a)) atomicAdd(reinterpret_cast<unsigned long long*>(a.g_fx) + a.sc0 + k, s[k]);
      else atomicAdd(a.g_acc + a.sc0 + k, *reinterpret_cast<double*>(&s[k]));
    }
  }
};
// dynamic shared memory of every tracking kernel ahead of the tally accumulators: block slots + per-thread slots
constexpr int COUNTER_SMEM_BYTES = RB_NSCALARS * 8 + TRACK_THREADS * 8 + 5 * TRACK_THREADS * 4 + 8 + TRACK_THREADS * 8;

// ======================================================================================
// Transport.MC — 1-D history-based tracking
// ======================================================================================
template <class P>
struct TrackArgs {
  MeshDev<P> m;
  Parts<P> p;
  long long n;
  typename P::comp_t dt;
  RngArgs rng;
  TallyArgs tally;
  signed char* out_event;  // optional per-particle outcome record (replay checks)
  int* out_nseg;
  unsigned long long* over_flag;
  // random-walk tables (MC_RW)
  const typename P::store_t *aVals, *ptVals;
  int n_rw_table;
  // event-based schedule: one segment per particle per launch
  const unsigned* ev_in;      // active particle indices (nullptr: 0..n-1)
  unsigned* ev_out;           // particles still in flight after this launch
  unsigned long long* ev_count;  // [0] = entries written to ev_out
  int* ev_nseg;               // [n] segments tracked so far
  unsigned* ev_extra;         // [n] extra-stream words consumed so far
  int ev_batch;               // segments per particle per launch
  // dynamic schedule
  unsigned long long* queue;  // next unclaimed chunk ticket
  unsigned long long queue_chunks;   // chunks of QUEUE_CHUNK particles in the list
  int refill_min;             // refill when at least this many lanes of a warp are idle
  // optional timeline of the dynamic schedule (IMC_TRACK_TIMING=1): [0] first block start, [1] first time a warp found the
  // queue empty, [2] last block exit (globaltimer ns)
  unsigned long long* timeline;
};

// One particle's registers while it is being tracked, and one loop iteration ("segment") as a device
// function, shared by the two history-based schedules:
//   k_track*         static: thread t takes particles t, t + stride, ...; a warp waits for its longest history
//   k_track_refill   dynamic: lanes whose history ended take the next particles from a global queue
//                    (warp-aggregated atomic) while the other lanes keep tracking — keeps lanes busy when
//                    history lengths differ by orders of magnitude (thin pipe vs hot thick wall)
// seg*() returns -1 to continue, or the outcome 0 census / 1 absorbed / 2 escaped.
template <class P>
struct Hist1 {
  Num<P> t, x, mu, E, E0, minE;
  int cell, k, nseg;
  unsigned pi;          // position in the particle list (< 2^32, checked by the engine)
  long long rec_base;
};
template <class P, class D, int TK>
__device__ __forceinline__ bool load1d(const TrackArgs<P>& a, long long pi, Hist1<P>& h, D& d, Counters& cn) {
  using N = Num<P>;
  h.E0 = N::load(a.p.E0, pi);
  if (h.E0.v == (typename P::comp_t)-1) return false;  // flagged dead and not yet cleaned
  h.pi = (unsigned)pi;
  h.t = N::load(a.p.t, pi); h.x = N::load(a.p.x, pi); h.mu = N::load(a.p.mu, pi); h.E = N::load(a.p.E, pi);
  h.cell = a.p.cx[pi];
  h.k = a.p.ks[pi];
  h.minE = N::from_d(0.01 * h.E0.d());                                              // :61
  h.nseg = 0;
  h.rec_base = (TKind<TK>::exact(a.tally) && a.tally.pass == 2) ? a.tally.rec_off[pi] : 0;
  d.init(a.rng, a.p.id[pi], pi);
  return true;
}
template <class P, class D, int TK>
__device__ __forceinline__ void store1d(const TrackArgs<P>& a, Hist1<P>& h, D& d, int ev, Counters& cn) {
  const long long pi = h.pi;
  if (TKind<TK>::exact(a.tally) && a.tally.pass == 1) { a.tally.rec_cnt[pi] = h.nseg; return; }
  cn.finish(ev, h.nseg);
  if (ev == 0) { h.t.store(a.p.t, pi); h.x.store(a.p.x, pi); h.mu.store(a.p.mu, pi); h.E.store(a.p.E, pi); a.p.cx[pi] = h.cell; }
  else h.E0.store(a.p.E0, pi);  // dead: only the flag is written; the other slots stay stale (Q16)
  if (a.out_event) { a.out_event[pi] = (signed char)ev; a.out_nseg[pi] = h.nseg; }
  if (d.over()) atomicAdd(a.over_flag, 1ull);
}
template <class P, class D, int TK>
__device__ __forceinline__ int seg1d(const TrackArgs<P>& a, Hist1<P>& h, D& d, Tally<P, TK>& tal, Counters& cn) {
  using N = Num<P>;
  const N one = N::from_d(1.0), two = N::from_i(2), zero;
  const N dt(a.dt), c_light(a.m.c), ds(a.m.ds);
  const int nc = (int)a.m.nc;
  ++h.nseg;                                                                         // :73
  const unsigned seg = (unsigned)h.nseg - 1u;
  d.next_segment(a.rng, seg);
  const CellProp1<P> cp = a.m.cp1[h.cell];
  const N w(P::unpack(cp.w)), dx(P::unpack(cp.dx)), sig_col(P::unpack(cp.sig_col)), neg_saf(P::unpack(cp.neg_saf));
  const bool fwd = h.mu > zero;
  const N qb = (fwd ? w - h.x : h.x) / h.mu;
  N dist_b = fwd ? qb : nabs(qb);                                                   // :77-83
  N dist_col = d.randexp(seg) / sig_col;                                            // :87
  N dist_cen = (c_light * (dt - h.t)) * ds;                                         // :89
  N dist = jl_min(jl_min(dist_b, dist_col), dist_cen);                              // :92
  N ex, em1; MathDet::exp_expm1<P>(neg_saf * dist, &ex, &em1);
  N newE = h.E * ex;                                                                // :95
  if (is_nan(newE) || is_nan(dist)) cn.error();
  const bool exact = TKind<TK>::exact(a.tally);
  const long long acc = (long long)h.k * nc + h.cell;
  // the deposit only feeds the tally: the reference's expression when the tally is reduced in reference order
  // (EXACT), else E * (1/dx) (<= 1.5 ulp from it; sums in these modes are order-dependent anyway)
  const N idx = exact ? N() : N(P::unpack(a.m.ax_inv[h.cell].q));
  if (newE <= h.minE) {                                                             // :97-106
    tal.add_runs((int)acc, exact ? h.E / dx : h.E * idx, h.rec_base + h.nseg - 1);
    h.E0 = N::from_d(-1.0);
    return 1;
  }
  tal.add_runs((int)acc, exact ? (-(h.E / dx)) * em1 : ((-h.E) * idx) * em1, h.rec_base + h.nseg - 1);                        // :110 / :120
  h.x = h.x + h.mu * dist;                                                          // :124
  { N dd = dist;                                                                    // :125 (dist / ds) / c; x / 1 == x exactly
#pragma unroll 1
    for (int r = 0; r < a.m.n_tdiv; ++r) dd = div_cached(dd, N(a.m.tdiv[r]), a.m.tdiv_r[r]);
    h.t = h.t + dd; }
  h.E = newE;                                                                       // :126
  bool dead = false;
  if (dist == dist_b) {                                                             // :130-170
    if (h.mu > zero) {
      if (h.cell == nc - 1) {
        if (a.m.bc[IMC_BC_RIGHT] == IMC_REFLECT) h.mu = -h.mu; else dead = true;
      }
      if (!dead) { h.cell += 1; h.x = zero; }
    }
    if (!dead && h.mu < zero) {
      if (h.cell == 0) {
        if (a.m.bc[IMC_BC_LEFT] == IMC_REFLECT) h.mu = -h.mu; else dead = true;
      } else { h.cell -= 1; h.x = N::load(a.m.wx, h.cell); }
    }
  }
  if (dead) { cn.lose<P, TK>(a.tally, h.E / N(a.m.scales[h.k]), h.pi, h.E.d()); h.E0 = N::from_d(-1.0); return 2; }  // :141 / :160
  if (dist == dist_col) {                                                           // :174-183
    h.mu = zero;
    while (h.mu == zero) h.mu = one - two * d.uniform(a.rng, seg);
  }
  if (dist == dist_cen) { h.t = zero; return 0; }                      // :185-193
  return -1;
}

template <class P, bool TAPE, int TK>
__global__ void __launch_bounds__(TRACK_THREADS, track_min_blocks<P>()) k_track1d(TrackArgs<P> a) {
  extern __shared__ __align__(16) unsigned char smem[];
  Counters cn; cn.init(reinterpret_cast<unsigned long long*>(smem));
  Tally<P, TK> tal(a.tally, smem + COUNTER_SMEM_BYTES);
  tal.zero();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x; pi < a.n; pi += stride) {
    Hist1<P> h; HistDraw<P, TAPE> d;
    if (!load1d<P, HistDraw<P, TAPE>, TK>(a, pi, h, d, cn)) continue;
    int ev;
    while ((ev = seg1d(a, h, d, tal, cn)) < 0) {}
    store1d<P, HistDraw<P, TAPE>, TK>(a, h, d, ev, cn);
  }
  tal.flush();
  cn.commit<TK>(a.tally);
}

// ======================================================================================
// Transport.MC2D — 2-D history-based tracking
// ======================================================================================
// qx / qy: 1/dx and 1/dy of the current cell — or dx and dy themselves when the deposits keep the reference's
// expression (E/dx)/dy (EXACT tallies).  The per-axis constants {d, w = d*ds, 1/d} of both axes sit in ONE table
// (x entries, then y entries), so a face crossing in either direction is the same code with a different index.
template <class P>
struct Hist2 {
  Num<P> t, x, y, mu, E, E0, minE, vx, vy, wxc, wyc, qx, qy, sig_col, neg_saf;
  typename P::comp_t rvx, rvy;   // cached reciprocals of vx, vy (they change only when mu does)
  int xi, yi, k, nseg;
  unsigned pi;          // position in the particle list (< 2^32, checked by the engine)
  long long rec_base;
};
// The direction angle mu is read once per history (at the census write-back; the loop works with cos / sin): Float64
// histories, which run at the register cap, keep it in the thread's shared-memory slot instead of two registers.
template <class P> __device__ __forceinline__ void set_mu(Hist2<P>& h, Counters& cn, Num<P> v) { if constexpr (P::id == 2) *cn.park = v.v; else h.mu = v; }
template <class P> __device__ __forceinline__ Num<P> get_mu(const Hist2<P>& h, const Counters& cn) { if constexpr (P::id == 2) return Num<P>(*cn.park); else return h.mu; }
template <class P, class D, int TK>
__device__ __forceinline__ bool load2d(const TrackArgs<P>& a, long long pi, Hist2<P>& h, D& d, Counters& cn) {
  using N = Num<P>;
  h.E = N::load(a.p.E, pi);
  if (h.E.v == (typename P::comp_t)-1) return false;  // 2-D dead flag lives in the energy slot (Q16)
  h.pi = (unsigned)pi;
  h.E0 = N::load(a.p.E0, pi);
  h.t = N::load(a.p.t, pi); h.x = N::load(a.p.x, pi); h.y = N::load(a.p.y, pi);
  const N mu0 = N::load(a.p.mu, pi); set_mu(h, cn, mu0);
  h.xi = a.p.cx[pi]; h.yi = a.p.cy[pi];
  h.k = a.p.ks[pi];
  h.minE = N::from_d(0.01 * h.E0.d());                                              // :531
  h.nseg = 0;
  const bool exact = TKind<TK>::exact(a.tally);
  h.rec_base = (exact && a.tally.pass == 2) ? a.tally.rec_off[pi] : 0;
  d.init(a.rng, a.p.id[pi], pi);
  MathDet::sincos<P>(mu0, &h.vy, &h.vx);                                            // :534 (recomputed only when mu changes)
  recip_dir(h.vx, h.vy, h.rvx, h.rvy);
  { const AxisProp<P>* tab = exact ? a.m.ax_d : a.m.ax_inv;
    const AxisProp<P> ax = tab[h.xi], ay = tab[a.m.nx + h.yi];
    h.wxc = N(P::unpack(ax.w)); h.wyc = N(P::unpack(ay.w)); h.qx = N(P::unpack(ax.q)); h.qy = N(P::unpack(ay.q)); }
  { const CellProp2<P> cp = load_cell2(a.m.cp2, h.xi * a.m.csx + h.yi * a.m.csy);
    h.sig_col = N(P::unpack(cp.sig_col)); h.neg_saf = N(P::unpack(cp.neg_saf)); }
  return true;
}
template <class P, class D, int TK>
__device__ __forceinline__ void store2d(const TrackArgs<P>& a, Hist2<P>& h, D& d, int ev, Counters& cn) {
  const long long pi = h.pi;
  if (TKind<TK>::exact(a.tally) && a.tally.pass == 1) { a.tally.rec_cnt[pi] = h.nseg; return; }
  cn.finish(ev, h.nseg);
  if (ev == 0) {
    h.t.store(a.p.t, pi); h.x.store(a.p.x, pi); h.y.store(a.p.y, pi); get_mu(h, cn).store(a.p.mu, pi); h.E.store(a.p.E, pi);
    a.p.cx[pi] = h.xi; a.p.cy[pi] = h.yi;
  } else h.E.store(a.p.E, pi);
  if (a.out_event) { a.out_event[pi] = (signed char)ev; a.out_nseg[pi] = h.nseg; }
  if (d.over()) atomicAdd(a.over_flag, 1ull);
}
// PAR: parity of the segment index (h.nseg before the increment) when the caller's loop is unrolled by two, else -1
template <int PAR = -1, class P, class D, int TK>
__device__ __forceinline__ int seg2d(const TrackArgs<P>& a, Hist2<P>& h, D& d, Tally<P, TK>& tal, Counters& cn) {
  using N = Num<P>;
  const N zero;
  const N dt(a.dt), c_light(a.m.c), ds(a.m.ds);
  const int nx = a.m.nx;
  const bool exact = TKind<TK>::exact(a.tally);
  ++h.nseg;
  const unsigned seg = (unsigned)h.nseg - 1u;
  d.template next_segment<PAR>(a.rng, seg);
  // the cell's {sigma_a (1 - f) + sigma_s, -sigma_a f} stay in registers and are re-read when the particle enters
  // another cell
  const N sig_col = h.sig_col, neg_saf = h.neg_saf;
  N qbx, qby;
  div_dir2<P>(h.vx > zero ? h.wxc - h.x : h.x, h.vx, h.rvx, h.vy > zero ? h.wyc - h.y : h.y, h.vy, h.rvy, qbx, qby);
  const N dist_bx = nabs(qbx);                                                      // :538-542
  const N dist_by = nabs(qby);                                                      // :544-548
  const N dist_b = min_nonnan(dist_bx, dist_by);                                    // :551-557
  const N dist_col = d.template randexp<PAR>(seg) / sig_col;                        // :561
  const N dist_cen = (c_light * (dt - h.t)) * ds;                                   // :569
  const N dist = jl_min(jl_min(dist_b, dist_col), dist_cen);                        // :571
  if (is_nan(dist) || dist_col < zero) cn.error();
  // The event follows from the distances alone (:621-622), so the constants of the cell the particle is about to enter
  // are requested NOW and used at the end of the segment: the gather latency (L2) is covered by the attenuation,
  // deposit and move below instead of stalling the next segment.  One code path for the four faces: axis = x if
  // dist_bx < dist_by (:622, false on NaN -> y), direction from the sign of that axis' direction cosine
  // (:623 / :643 / :663 / :683).
  const bool face = dist == dist_bx || dist == dist_by;                             // :621
  const bool isx = dist_bx < dist_by;
  const bool pos = (isx ? h.vx : h.vy) > zero;
  const int idx = isx ? h.xi : h.yi;
  const bool interior = face && (pos ? idx != (isx ? nx : a.m.ny) - 1 : idx != 0);
  const int ni = pos ? idx + 1 : idx - 1;
  const int nxi = isx ? ni : h.xi, nyi = isx ? h.yi : ni;
  const long long acc = (long long)h.k * a.m.nc + (h.xi * a.m.tsx + h.yi * a.m.tsy);   // nc < 2^31 (checked at set_mesh)
  AxisProp<P> nax; CellProp2<P> ncp;
  nax.w = nax.q = ncp.sig_col = ncp.neg_saf = typename P::store_t(0);
#if IMC_EARLY_NEXT_CELL
  if (interior) {
    nax = (exact ? a.m.ax_d : a.m.ax_inv)[(isx ? 0 : nx) + ni];
    ncp = load_cell2(a.m.cp2, nxi * a.m.csx + nyi * a.m.csy);
  }
#endif
  N ex, em1; MathDet::exp_expm1<P>(neg_saf * dist, &ex, &em1);
  const N newE = h.E * ex;                                                          // :580
  // the deposit only feeds the tally: the reference's expression when the tally is reduced in reference order
  // (EXACT), else E * ((1/dx) * (1/dy)) (<= 3 ulp from it; sums in these modes are order-dependent anyway)
  if (newE <= h.minE) {                                                             // :586-595
    tal.add_runs((int)acc, exact ? (h.E / h.qx) / h.qy : h.E * (h.qx * h.qy), h.rec_base + h.nseg - 1);
    h.E = N::from_d(-1.0);
    return 1;
  }
  tal.add_runs((int)acc, exact ? ((-(h.E / h.qx)) / h.qy) * em1 : ((-h.E) * (h.qx * h.qy)) * em1, h.rec_base + h.nseg - 1);  // :599 / :607
  h.x = h.x + dist * h.vx;                                                          // :615
  h.y = h.y + dist * h.vy;                                                          // :616
  { N dd = dist;                                                                    // :617 (dist / ds) / c; x / 1 == x exactly
    if (a.m.n_tdiv != 0) {                                                          // kernel-uniform branches
      dd = div_cached(dd, N(a.m.tdiv[0]), a.m.tdiv_r[0]);
      if (a.m.n_tdiv == 2) dd = div_cached(dd, N(a.m.tdiv[1]), a.m.tdiv_r[1]);
    }
    h.t = h.t + dd; }
  h.E = newE;                                                                       // :618
  if (face) {
    if (interior) {                                                                 // neighbour cell
#if !IMC_EARLY_NEXT_CELL
      nax = (exact ? a.m.ax_d : a.m.ax_inv)[(isx ? 0 : nx) + ni];
      ncp = load_cell2(a.m.cp2, nxi * a.m.csx + nyi * a.m.csy);
#endif
      const N w(P::unpack(nax.w)), q(P::unpack(nax.q));
      const N np = pos ? zero : w;                                                  // enters at 0 or at the far edge dx*ds
      h.xi = nxi; h.yi = nyi;
      h.sig_col = N(P::unpack(ncp.sig_col)); h.neg_saf = N(P::unpack(ncp.neg_saf));
      h.x = isx ? np : h.x; h.wxc = isx ? w : h.wxc; h.qx = isx ? q : h.qx;
      h.y = isx ? h.y : np; h.wyc = isx ? h.wyc : w; h.qy = isx ? h.qy : q;
      return -1;                                                                    // `continue` :703 (Q15)
    }
    const int side = isx ? (pos ? IMC_BC_RIGHT : IMC_BC_LEFT) : (pos ? IMC_BC_TOP : IMC_BC_BOTTOM);
    if (a.m.bc[side] == IMC_REFLECT) {                                              // :626-628 / :666-668
      const N mu1 = isx ? MathDet::atan2<P>(h.vy, -h.vx) : MathDet::atan2<P>(-h.vy, h.vx);
      set_mu(h, cn, mu1);
      MathDet::sincos<P>(mu1, &h.vy, &h.vx);
      recip_dir(h.vx, h.vy, h.rvx, h.rvy);
      return -1;
    }
    cn.lose<P, TK>(a.tally, h.E / N(a.m.scales[h.k]), h.pi, h.E.d());               // VACUUM :629-636 ...
    h.E = N::from_d(-1.0);
    return 2;
  }
  if (dist == dist_col) { const N mu1 = N::from_d(6.283185307179586 * d.template uniform<PAR>(a.rng, seg).d()); set_mu(h, cn, mu1); MathDet::sincos<P>(mu1, &h.vy, &h.vx);
                           recip_dir(h.vx, h.vy, h.rvx, h.rvy); }  // :706-710
  if (dist == dist_cen) { h.t = zero; return 0; }                      // :712-717
  return -1;
}

template <class P, bool TAPE, int TK>
__global__ void __launch_bounds__(TRACK_THREADS, track_min_blocks<P>()) k_track2d(TrackArgs<P> a) {
  extern __shared__ __align__(16) unsigned char smem[];
  Counters cn; cn.init(reinterpret_cast<unsigned long long*>(smem));
  Tally<P, TK> tal(a.tally, smem + COUNTER_SMEM_BYTES);
  tal.zero();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x; pi < a.n; pi += stride) {
    Hist2<P> h; HistDraw<P, TAPE, true> d;
    if (!load2d<P, HistDraw<P, TAPE, true>, TK>(a, pi, h, d, cn)) continue;
    int ev;   // two segments per trip: the parity of the segment index is static (one Philox block per pair)
    while ((ev = seg2d<0>(a, h, d, tal, cn)) < 0 && (ev = seg2d<1>(a, h, d, tal, cn)) < 0) {}
    store2d<P, HistDraw<P, TAPE, true>, TK>(a, h, d, ev, cn);
  }
  tal.flush();
  cn.commit<TK>(a.tally);
}

// ---- event-based schedule: every launch advances each in-flight particle by up to `ev_batch` segments ----------
// A launch loads the state of every particle still in flight, tracks it until it ends or has done ev_batch segments,
// writes the state of the unfinished ones back and appends them to the next launch's index list with one warp-aggregated
// atomic (stream compaction by event: the next launch starts with full warps again).  The price against the
// history schedules is B_hist per ev_batch segments instead of per history, the direction vector and the Philox block
// recomputed at every reload, and one launch + one count read-back per batch; ev_batch = 1 is the textbook
// one-segment-per-launch form (17-40x slower here: the longest histories need thousands of nearly empty launches).
// Philox only (the replay tape has per-particle cursors); EXACT tallies use the history schedules.
template <class P, int GEOM, int TK>
__global__ void __launch_bounds__(TRACK_THREADS, track_min_blocks<P>()) k_track_event(TrackArgs<P> a, long long n_active, int first) {
  extern __shared__ __align__(16) unsigned char smem[];
  Counters cn; cn.init(reinterpret_cast<unsigned long long*>(smem));
  Tally<P, TK> tal(a.tally, smem + COUNTER_SMEM_BYTES);
  tal.zero();
  using Dr = HistDraw<P, false>;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_round = (n_active + 31) & ~31ll;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    bool cont = false;
    long long pi = 0;
    if (i < n_active) {
      pi = a.ev_in ? (long long)a.ev_in[i] : i;
      Hist1<P> h1; Hist2<P> h2; Dr d;
      bool ok;
      if constexpr (GEOM == 1) ok = load1d<P, Dr, TK>(a, pi, h1, d, cn); else ok = load2d<P, Dr, TK>(a, pi, h2, d, cn);
      if (ok) {
        const int done = first ? 0 : a.ev_nseg[pi];
        if (!first) d.resume(a.p.id[pi], (unsigned)done, a.ev_extra[pi]);
        int ev = -1;
        if constexpr (GEOM == 1) {
          h1.nseg = done;
          for (int k = 0; k < a.ev_batch && ev < 0; ++k) ev = seg1d(a, h1, d, tal, cn);
          if (ev >= 0) store1d<P, Dr, TK>(a, h1, d, ev, cn);
          else { h1.t.store(a.p.t, pi); h1.x.store(a.p.x, pi); h1.mu.store(a.p.mu, pi); h1.E.store(a.p.E, pi); a.p.cx[pi] = h1.cell; a.ev_nseg[pi] = h1.nseg; }
        } else {
          h2.nseg = done;
          for (int k = 0; k < a.ev_batch && ev < 0; ++k) ev = seg2d(a, h2, d, tal, cn);
          if (ev >= 0) store2d<P, Dr, TK>(a, h2, d, ev, cn);
          else { h2.t.store(a.p.t, pi); h2.x.store(a.p.x, pi); h2.y.store(a.p.y, pi); get_mu(h2, cn).store(a.p.mu, pi); h2.E.store(a.p.E, pi);
                 a.p.cx[pi] = h2.xi; a.p.cy[pi] = h2.yi; a.ev_nseg[pi] = h2.nseg; }
        }
        if (ev < 0) { a.ev_extra[pi] = d.sg.extra_n & 0x3fffffffu; cont = true; }
      }
    }
    const unsigned m = __ballot_sync(IMC_FULL_MASK, cont);
    if (m) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.ev_count, (unsigned long long)__popc(m));
      base = __shfl_sync(IMC_FULL_MASK, base, 0);
      if (cont) a.ev_out[base + __popc(m & ((1u << lane) - 1u))] = (unsigned)pi;
    }
  }
  tal.flush();
  cn.commit<TK>(a.tally);
}

// ======================================================================================
// Transport.MC_RW — 1-D tracking with random-walk acceleration
// ======================================================================================
// A value whose Julia type is T or Float64 at run time: MC_RW draws randexp() in Float64 and ignores
// the distance scale, so position / time / energy are promoted to Float64 inside a history after the
// first ordinary segment and only rounded to T on write-back (Q3).
template <class P>
struct DynD {
  double v; bool wide;
  __device__ __forceinline__ DynD() : v(0), wide(false) {}
  __device__ __forceinline__ DynD(Num<P> x) : v(x.d()), wide(false) {}
  static __device__ __forceinline__ DynD w(double x) { DynD r; r.v = x; r.wide = true; return r; }
  __device__ __forceinline__ Num<P> narrow() const { return Num<P>::from_d(v); }
};
// One code path for both cases (the lanes of a warp hold first-segment histories, still of type T, next to later ones that are
// Float64 already): the operation is done in Float64 and the result rounded to T when both operands are T.  For + - * / of two
// T values that equals T's own operation — a binary64 intermediate has more than 2p + 2 digits for p = 24 (and Float16 operations
// are Float32 operations rounded once more, as in Julia), so the second rounding is innocuous.
#define IMC_DYN_OP(name, op)                                                                         \
  template <class P> __device__ __forceinline__ DynD<P> name(DynD<P> a, DynD<P> b) {                \
    DynD<P> r; r.wide = a.wide || b.wide;                                                           \
    const double v = a.v op b.v;                                                                    \
    if constexpr (P::id == 2) r.v = v; else r.v = r.wide ? v : (double)P::rnd((float)v);            \
    return r;                                                                                       \
  }
IMC_DYN_OP(dyn_add, +)
IMC_DYN_OP(dyn_sub, -)
IMC_DYN_OP(dyn_mul, *)
IMC_DYN_OP(dyn_div, /)
#undef IMC_DYN_OP
template <class P> __device__ __forceinline__ DynD<P> dyn_abs(DynD<P> a) { a.v = a.v < 0 ? -a.v : (a.v == 0 ? 0.0 : a.v); return a; }
template <class P> __device__ __forceinline__ DynD<P> dyn_min(DynD<P> a, DynD<P> b) {
  DynD<P> r; r.wide = a.wide || b.wide; r.v = jl_min(Num<F64>(a.v), Num<F64>(b.v)).v; return r;
}

// Transport.P_r (imc_transport.jl:734-754): 100-term series accumulated in Float64 (Q31).  Terms are
// added in order; once exp() has underflowed to exactly 0 the remaining terms cannot change the sum.
__device__ __forceinline__ double rw_P_r(double a) {
  if (a == 0) return 1.0;
  double Pr = 0.0;
  for (int n = 1; n <= 100; ++n) {
    double pin = 3.141592653589793 * (double)n;
    double e = dm::exp_d(-a * (pin * pin));
    // The terms decrease monotonically (a > 0), so once 2e is below a quarter of the last place of the running sum no
    // later term can change it (round to nearest, |term| < ulp / 2 strictly): the remaining iterations of the reference's
    // loop are no-ops and are skipped.  (e == 0: underflow, covers Pr == 0 as well.)  a < 0 or NaN: the test never fires.
    if (e == 0.0 || e * 2.0 < (Pr < 0 ? -Pr : Pr) * 0x1p-55) break;
    Pr += (((n - 1) & 1) ? -1.0 : 1.0) * e * 2.0;
  }
  return Pr;
}
// Transport.bisection (imc_transport.jl:756-784), 1-based
template <class P>
__device__ __forceinline__ int rw_bisection(const typename P::store_t* arr, int n, double value) {
  if (value < Num<P>::load(arr, 0).d()) return 1;
  if (value > Num<P>::load(arr, n - 1).d()) return n;
  int jl = 1, ju = n;
  while (ju - jl > 1) {
    int jm = (ju + jl) >> 1;
    if (value >= Num<P>::load(arr, jm - 1).d()) jl = jm; else ju = jm;
  }
  if (value == Num<P>::load(arr, 0).d()) return 1;
  if (value == Num<P>::load(arr, n - 1).d()) return n;
  return jl;
}

// Draw source of the MC_RW loops: words are consumed in sequence (a segment draws one Float64 exponential, a random-walk
// trial one or two uniforms, a collision one or more), so there is a cursor — the Philox variant keeps only what differs
// per thread (particle id, block index, the four buffered words); the key and the step sit in RngArgs.  Same words as
// PhiloxDraw<P> on (seed, id, step, STREAM_TRACK), which the oracle uses.
template <class P, bool TAPE> struct RwDraw;
template <class P>
struct RwDraw<P, false> {
  uint32_t id_lo, id_hi, blk, used, buf[4];
  __device__ __forceinline__ void init(const RngArgs&, unsigned long long id, long long) { id_lo = (uint32_t)id; id_hi = (uint32_t)(id >> 32); blk = 0u; used = 4u; buf[0] = buf[1] = buf[2] = buf[3] = 0u; }
  __device__ __forceinline__ uint32_t word(const RngArgs& r) {
    if (used == 4u) { const uint32_t c[4] = {id_lo, id_hi, r.step | (STREAM_TRACK << 28), blk}; Philox::block_rk(c, r.rk, buf); ++blk; used = 0u; }
    const uint32_t w = used == 0u ? buf[0] : used == 1u ? buf[1] : used == 2u ? buf[2] : buf[3];
    ++used;
    return w;
  }
  __device__ __forceinline__ Num<P> uniform(const RngArgs& r) {
    if constexpr (P::id == 2) { const uint64_t lo = word(r), hi = word(r); return uniform_from_word<P>((hi << 32) | lo); }
    else return uniform_from_word<P>((uint64_t)word(r));
  }
  __device__ __forceinline__ double randexp64(const RngArgs& r) { const uint64_t lo = word(r), hi = word(r); return randexp64_from_word((hi << 32) | lo); }
  __device__ __forceinline__ bool over() const { return false; }
};
template <class P>
struct RwDraw<P, true> {
  TapeDraw<P> tp;
  __device__ __forceinline__ void init(const RngArgs& r, unsigned long long, long long slot) { tp.init(r.uni, r.n_uni, r.ex, r.n_exp, (size_t)r.stride, (size_t)slot); }
  __device__ __forceinline__ Num<P> uniform(const RngArgs&) { return tp.uniform(); }
  __device__ __forceinline__ double randexp64(const RngArgs&) { return tp.randexp64(); }
  __device__ __forceinline__ bool over() const { return tp.exhausted(); }
};

// One MC_RW history in registers, and one loop iteration as a device function shared by the static and the warp-refill
// schedule.  Returns -1 to continue or the outcome: 0 census, 1 absorbed, 2 escaped, 3 killed by a random-walk step (Q1).
template <class P>
struct HistRW {
  DynD<P> t, x, E;      // values of T until the history's first move, Float64 afterwards (Q3)
  Num<P> mu, E0, minE;
  int cell, k, nseg;
  unsigned pi;
  long long rec_base;
};
template <class P, class D, int TK>
__device__ __forceinline__ bool load_rw(const TrackArgs<P>& a, long long pi, HistRW<P>& h, D& d, Counters&) {
  using N = Num<P>;
  h.E0 = N::load(a.p.E0, pi);
  if (h.E0.v == (typename P::comp_t)-1) return false;
  h.pi = (unsigned)pi;
  h.t = DynD<P>(N::load(a.p.t, pi)); h.x = DynD<P>(N::load(a.p.x, pi)); h.E = DynD<P>(N::load(a.p.E, pi));
  h.mu = N::load(a.p.mu, pi);
  h.cell = a.p.cx[pi]; h.k = a.p.ks[pi];
  h.minE = N::from_d(0.01 * h.E0.d());                                              // :262
  h.nseg = 0;
  h.rec_base = (TKind<TK>::exact(a.tally) && a.tally.pass == 2) ? a.tally.rec_off[pi] : 0;
  d.init(a.rng, a.p.id[pi], pi);
  return true;
}
template <class P, class D, int TK>
__device__ __forceinline__ void store_rw(const TrackArgs<P>& a, HistRW<P>& h, D& d, int ev, Counters& cn) {
  using N = Num<P>;
  const long long pi = h.pi;
  if (TKind<TK>::exact(a.tally) && a.tally.pass == 1) { a.tally.rec_cnt[pi] = h.nseg; return; }
  cn.finish(ev, h.nseg);
  if (ev == 0) {
    N().store(a.p.t, pi); N::from_d(h.x.v).store(a.p.x, pi); h.mu.store(a.p.mu, pi); N::from_d(h.E.v).store(a.p.E, pi); a.p.cx[pi] = h.cell;
  } else h.E0.store(a.p.E0, pi);
  if (a.out_event) { a.out_event[pi] = (signed char)ev; a.out_nseg[pi] = h.nseg; }
  if (d.over()) atomicAdd(a.over_flag, 1ull);
}
template <class P, class Dr, int TK>
__device__ __forceinline__ int seg_rw(const TrackArgs<P>& a, HistRW<P>& h, Dr& d, Tally<P, TK>& tal, Counters& cn) {
  using N = Num<P>;
  using D = DynD<P>;
  const N one = N::from_d(1.0), two = N::from_i(2), three = N::from_i(3), zero;
  const N dt(a.dt), c_light(a.m.c);
  const int nc = (int)a.m.nc;
  const bool exact = TKind<TK>::exact(a.tally);
  const long long acc = (long long)nc * h.k + h.cell;
  ++h.nseg;                                                                         // :265
  const CellProp1<P> cp = a.m.cp1[h.cell];
  const N dx(P::unpack(cp.dx)), sig_col(P::unpack(cp.sig_col)), neg_saf(P::unpack(cp.neg_saf));
  const bool fwd = h.mu > zero;
  const D qb = dyn_div(fwd ? dyn_sub(D(dx), h.x) : h.x, D(h.mu));
  const D dist_b = fwd ? qb : dyn_abs(qb);                                          // :269-275 (no distancescale)
  const double rex = d.randexp64(a.rng);
  const D dist_col = D::w((rex < 0 ? -rex : rex) / sig_col.d());                    // :279 (Float64)
  const D dist_cen = dyn_mul(D(c_light), dyn_sub(D(dt), h.t));                      // :281
  const D dist = dyn_min(dyn_min(dist_b, dist_col), dist_cen);                      // :284
  const D R0 = dyn_min(dyn_abs(dyn_sub(D(dx), h.x)), dyn_abs(h.x));                 // :286
  const N inv_sigma = N::from_i(1) / N::load(a.m.sigma_static, h.cell);             // 1/mesh.sigma[cellindex] (Q4)
  if (R0.v > inv_sigma.d() && dist_col.v < R0.v) {                                  // :289
    const N sa = N::load(a.m.sa, h.cell), f = N::load(a.m.fleck, h.cell);
    const N u = d.uniform(a.rng);                                                   // :290
    const N Dc = c_light / ((three * sa) * (one - f));                              // :292
    const D aa = dyn_div(dyn_mul(D(Dc), D(dt)), dyn_mul(R0, R0));                   // :294
    const double Pr = rw_P_r(aa.v);                                                 // :296
    const double Pt = 1.0 - Pr;                                                     // :297
    const N lg = MathDet::log<P>(one - f);
    N expo;
    if (u.d() < Pt) {                                                               // :298
      const int ai = rw_bisection<P>(a.ptVals, a.n_rw_table, u.d());                // :301
      const D tp = dyn_div(dyn_mul(D(N::load(a.aVals, ai - 1)), dyn_mul(R0, R0)), D(Dc));
      const N t_p = N::from_d(tp.v);                                                // :303
      expo = (((t_p * c_light) * (one - f)) * sa) / lg;                             // :306
    } else {
      (void)d.uniform(a.rng);                                                       // u_prime :338
      expo = (((c_light * (one - f)) * sa) * dt) / lg;                              // :346
    }
    // newenergy <= startenergy always holds (Q1): the random-walk step deposits and kills
    N ex, em1; MathDet::exp_expm1<P>(expo, &ex, &em1);
    const D newE = dyn_mul(h.E, D(ex));
    const D depv = dyn_mul(dyn_mul(D(-one), dyn_div(h.E, D(dx))), D(em1));          // :317-320 / :352-356
    if (exact) tal.add(acc, N::from_d(depv.v), h.rec_base + h.nseg - 1, depv.wide, depv.v);   // Float64 deposits: converted on push! / added in Float64 on setindex!
    else tal.add_runs((int)acc, N::from_d(depv.v));
    if (newE.v != newE.v) cn.error();
    h.E0 = N::from_d(-1.0);
    return 3;
  }
  D newE = dyn_mul(h.E, D::w(dm::exp_d(neg_saf.d() * dist.v)));                     // :376 (Float64)
  if (newE.v <= h.minE.d()) newE = D(zero);                                         // :377-379
  const D depv = dyn_sub(h.E, newE);                                                // :383 / :385 (not / dx, Q2)
  if (exact) tal.add(acc, N::from_d(depv.v), h.rec_base + h.nseg - 1, depv.wide, depv.v);
  else tal.add_runs((int)acc, N::from_d(depv.v));
  if (newE.v == 0.0) { h.E0 = N::from_d(-1.0); return 1; }                          // :390-394
  h.x = dyn_add(h.x, dyn_mul(D(h.mu), dist));                                       // :397
  h.t = dyn_add(h.t, dyn_div(dist, D(c_light)));                                    // :398
  h.E = newE;                                                                       // :399
  bool dead = false;
  if (dist.v == dist_b.v) {                                                         // :403-443
    if (h.mu > zero) {
      if (h.cell == nc - 1) { if (a.m.bc[IMC_BC_RIGHT] == IMC_REFLECT) h.mu = -h.mu; else dead = true; }
      if (!dead) { h.cell += 1; h.x = D(zero); }
    }
    if (!dead && h.mu < zero) {
      if (h.cell == 0) { if (a.m.bc[IMC_BC_LEFT] == IMC_REFLECT) h.mu = -h.mu; else dead = true; }
      else { h.cell -= 1; h.x = D(N::load(a.m.dx, h.cell)); }
    }
  }
  if (dead) {                                                                       // :414 / :433
    const N escale(a.m.scales[h.k]);
    if (exact) { if (a.tally.pass == 2) a.tally.lost_val[h.pi] = h.E.v; }
    else if (h.E.wide) cn.lose_value<TK>(a.tally, h.E.v / escale.d());
    else cn.lose<P, TK>(a.tally, h.E.narrow() / escale);
    h.E0 = N::from_d(-1.0);
    return 2;
  }
  if (dist.v == dist_col.v) {                                                       // :446-453
    h.mu = one - two * d.uniform(a.rng);
    while (h.mu == zero) h.mu = one - two * d.uniform(a.rng);
  }
  if (dist.v == dist_cen.v) return 0;                                               // :455-463
  return -1;
}

// static schedule (thread t takes particles t, t + stride, ...)
template <class P, bool TAPE, int TK>
__global__ void __launch_bounds__(TRACK_THREADS) k_track1d_rw(TrackArgs<P> a) {
  extern __shared__ __align__(16) unsigned char smem[];
  Counters cn; cn.init(reinterpret_cast<unsigned long long*>(smem));
  Tally<P, TK> tal(a.tally, smem + COUNTER_SMEM_BYTES);
  tal.zero();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x; pi < a.n; pi += stride) {
    HistRW<P> h; RwDraw<P, TAPE> d;
    if (!load_rw<P, RwDraw<P, TAPE>, TK>(a, pi, h, d, cn)) continue;
    int ev;
    while ((ev = seg_rw(a, h, d, tal, cn)) < 0) {}
    store_rw<P, RwDraw<P, TAPE>, TK>(a, h, d, ev, cn);
  }
  tal.flush();
  cn.commit<TK>(a.tally);
}

// ---- dynamic schedule: warp-level refill from a global particle queue --------------------------------
// Every trip of the loop each active lane tracks TWO segments (an even and an odd one: the parity of the segment index
// is static, so the Philox block of the pair is generated in the first half by every lane and the second half carries
// no parity test).  Before a trip, when at least `refill_min` lanes of the warp are idle (or all are), the idle lanes
// write back the histories they finished (thread-private counters, coalescing stores), lane 0 claims that many
// consecutive particle indices with one atomicAdd and the idle lanes load them.  Per-particle results do not depend on
// the lane that tracks them (Philox is keyed by particle id and segment, the tape by particle slot), so both schedules
// give identical particle state.
// GEOM: 1 = MC (1-D), 2 = MC2D, 3 = MC_RW (1-D with random-walk acceleration; its segments are Float64-heavy and carry no
// parity, so one per trip)
template <class P, int GEOM, bool TAPE> struct RefillDraw { using type = HistDraw<P, TAPE, GEOM == 2>; };
template <class P, bool TAPE> struct RefillDraw<P, 3, TAPE> { using type = RwDraw<P, TAPE>; };
#ifndef IMC_RW_MIN_BLOCKS
#define IMC_RW_MIN_BLOCKS 4   // MC_RW under the refill schedule: 4 blocks per SM at 64 registers (32 B of spill) measured 3 % faster than 3 at 80
#endif
template <class P, int GEOM, bool TAPE, int TK>
__global__ void __launch_bounds__(TRACK_THREADS, (GEOM == 3 ? IMC_RW_MIN_BLOCKS : track_min_blocks<P>())) k_track_refill(TrackArgs<P> a) {
  extern __shared__ __align__(16) unsigned char smem[];
  Counters cn; cn.init(reinterpret_cast<unsigned long long*>(smem));
  Tally<P, TK> tal(a.tally, smem + COUNTER_SMEM_BYTES);
  tal.zero();
  using Dr = typename RefillDraw<P, GEOM, TAPE>::type;
  constexpr int ST_EMPTY = -2, ST_ACTIVE = -1;   // >= 0: history finished with that outcome, not yet written back
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  int st = ST_EMPTY;
  bool drained = false;
  long long cbase = 0; int crem = 0;   // the warp's chunk of the particle list: next index, particles left (warp-uniform)
  unsigned long long tnext = 0;        // lane 0: the ticket claimed ahead
  if (lane == 0) tnext = atomicAdd(a.queue, 1ull);
  if (a.timeline && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); atomicMin(a.timeline, t); }
  Hist1<P> h1; Hist2<P> h2; HistRW<P> hr; Dr d;   // the two that GEOM does not use are never touched
  while (true) {
    unsigned idle = __ballot_sync(IMC_FULL_MASK, st != ST_ACTIVE);
    // refill_min <= 32, so a warp with no active lane always refills
    if (__popc(idle) >= a.refill_min) {
      if (st >= 0) {
        if constexpr (GEOM == 1) store1d<P, Dr, TK>(a, h1, d, st, cn); else if constexpr (GEOM == 2) store2d<P, Dr, TK>(a, h2, d, st, cn); else store_rw<P, Dr, TK>(a, hr, d, st, cn);
        st = ST_EMPTY;
      }
      if (!drained) {
        // the warp works through a private chunk of QUEUE_CHUNK consecutive particles (coalesced loads) and claims the
        // next one with a ticket from the global queue.  The ticket for the chunk AFTER the next is requested at once
        // (lane 0 keeps it), so the latency of the atomic is covered by a whole chunk of tracking.
        if (crem == 0) {
          unsigned long long t = tnext;
          if (lane == 0) tnext = atomicAdd(a.queue, 1ull);
          t = __shfl_sync(IMC_FULL_MASK, t, 0);
          if (t >= a.queue_chunks) {
            if (a.timeline && lane == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tt)); atomicMin(a.timeline + 1, tt); }
            drained = true;
          } else {
            cbase = (long long)t * QUEUE_CHUNK;
            crem = (int)min((long long)QUEUE_CHUNK, a.n - cbase);
          }
        }
        if (!drained) {
          const int take = min(__popc(idle), crem);
          const int rank = __popc(idle & lt_mask);
          if (st != ST_ACTIVE && rank < take) {
            const long long pi = cbase + rank;
            bool ok;
            if constexpr (GEOM == 1) ok = load1d<P, Dr, TK>(a, pi, h1, d, cn); else if constexpr (GEOM == 2) ok = load2d<P, Dr, TK>(a, pi, h2, d, cn); else ok = load_rw<P, Dr, TK>(a, pi, hr, d, cn);
            if (ok) st = ST_ACTIVE;
          }
          cbase += take; crem -= take;
        }
        idle = __ballot_sync(IMC_FULL_MASK, st != ST_ACTIVE);
      }
      if (idle == IMC_FULL_MASK) {
        if (drained) break;
        continue;
      }
    }
    if (st == ST_ACTIVE) {
      int ev;
      if constexpr (GEOM == 1) ev = seg1d(a, h1, d, tal, cn); else if constexpr (GEOM == 2) ev = seg2d<0>(a, h2, d, tal, cn); else ev = seg_rw(a, hr, d, tal, cn);
      if (ev >= 0) st = ev;
    }
    if constexpr ((P::id != 2 || IMC_UNROLL2_F64) && GEOM != 3) {
    if (st == ST_ACTIVE) {
      int ev;
      if constexpr (GEOM == 1) ev = seg1d(a, h1, d, tal, cn); else ev = seg2d<1>(a, h2, d, tal, cn);
      if (ev >= 0) st = ev;
    }
    }
  }
  if (a.timeline && lane == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); atomicMax(a.timeline + 2, t); }
  tal.flush();
  cn.commit<TK>(a.tally);
}


// ======================================================================================
// Sourcing.sample_planck (imc_sourcing.jl:372-399) — thread per sample.  Unused by the reference's step (every call
// site is commented out); exported through imc_sample_planck.
// ======================================================================================
constexpr int PLANCK_MAX_TERMS = 100000;   // the reference's loop has no exit when rn1 > max(90 nsum / pi^4): NaN by convention
template <class P>
__global__ void k_sample_planck(RngArgs r, long long n, double* __restrict__ out, unsigned long long* over_flag) {
  using N = Num<P>;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Draw<P> d; d.init(r, (unsigned long long)i, STREAM_PLANCK, i);
  const N one = N::from_d(1.0);
  const double pi2 = 3.141592653589793 * 3.141592653589793, pi4 = pi2 * pi2;        // pi^4: power by squaring in Float64
  N nn = one, nsum = one;                                                           // :382, :384
  N rn1 = d.uniform();                                                              // :383
  double res = nan("");
  for (int it = 0; it < PLANCK_MAX_TERMS; ++it) {
    if (rn1.d() <= (90.0 * nsum.d()) / pi4) {                                       // :388 (Float64 comparison)
      rn1 = d.uniform();                                                            // :389-392
      const N rn2 = d.uniform(), rn3 = d.uniform(), rn4 = d.uniform();
      const N lg = MathDet::log<P>(((rn1 * rn2) * rn3) * rn4);
      res = N::from_d((-1.0 * lg.d()) / nn.d()).d();                                // :393
      break;
    }
    nn = nn + one;                                                                  // :396
    // n^4: Float16 computes Float32(n)^4 and rounds once; n is a small integer, so every algorithm is exact in Float32 / Float64
    const double n4 = P::id == 0 ? N::from_d((double)nn.v * nn.v * nn.v * nn.v).d() : ((nn * nn) * (nn * nn)).d();
    nsum = nsum + N::from_d(1.0 / n4);                                              // :397
  }
  out[i] = res;
  if (d.over()) atomicAdd(over_flag, 1ull);
}

// ======================================================================================
// Clean.clean — stable compaction
// ======================================================================================
#ifndef IMC_COMPACT_MIN_BLOCKS
#define IMC_COMPACT_MIN_BLOCKS 4
#endif
constexpr int COMPACT_THREADS = 512;
constexpr int COMPACT_ITEMS = 8;                                // sub-tiles of COMPACT_THREADS particles per block
constexpr int COMPACT_TILE = COMPACT_THREADS * COMPACT_ITEMS;   // particles per block: one count per 4096 particles keeps the
                                                                // single-block scan of the counts short (24 k entries for 10^8 particles)

template <class P>
__device__ __forceinline__ bool particle_alive(const Parts<P>& p, long long i, int geom) {
  // slot 8 == -1.0: startenergy in 1-D, energy in 2-D (imc_clean.jl:15, Q16)
  return geom == 1 ? (P::unpack(p.E0[i]) != (typename P::comp_t)-1) : (P::unpack(p.E[i]) != (typename P::comp_t)-1);
}

// survivors per block of COMPACT_TILE particles: the eight flag loads of a thread are independent and coalesced
template <class P>
__global__ void __launch_bounds__(COMPACT_THREADS) k_alive_count(Parts<P> p, long long n, int geom, long long* __restrict__ block_counts) {
  __shared__ int warp_cnt[COMPACT_THREADS / 32];
  const long long base = (long long)blockIdx.x * COMPACT_TILE + threadIdx.x;
  bool alive[COMPACT_ITEMS];
#pragma unroll
  for (int k = 0; k < COMPACT_ITEMS; ++k) { const long long i = base + (long long)k * COMPACT_THREADS; alive[k] = i < n && particle_alive(p, i, geom); }
  int c = 0;
#pragma unroll
  for (int k = 0; k < COMPACT_ITEMS; ++k) c += __popc(__ballot_sync(IMC_FULL_MASK, alive[k]));
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < COMPACT_THREADS / 32; ++w) t += warp_cnt[w];
    block_counts[blockIdx.x] = t;
  }
}

// stable: sub-tile by sub-tile, inside a sub-tile by warp, inside a warp by lane.  The eight flag loads of a thread are in
// flight together, every (sub-tile, warp) count goes to shared memory behind ONE barrier, and the copies of the survivors
// follow with all destinations known (with 7 % survivors, as on the crooked pipe, the kernel is a chain of dependent memory
// round trips: one sub-tile per barrier made it eight chains long).
template <class P>
__global__ void __launch_bounds__(COMPACT_THREADS, IMC_COMPACT_MIN_BLOCKS) k_compact(Parts<P> src, Parts<P> dst, long long n, int geom, const long long* __restrict__ block_offs) {
  constexpr int WARPS = COMPACT_THREADS / 32;
  __shared__ int warp_cnt[COMPACT_ITEMS][WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long i0 = (long long)blockIdx.x * COMPACT_TILE + threadIdx.x;
  unsigned alive_mask = 0u;
#pragma unroll
  for (int k = 0; k < COMPACT_ITEMS; ++k) { const long long i = i0 + (long long)k * COMPACT_THREADS; alive_mask |= (i < n && particle_alive(src, i, geom)) ? (1u << k) : 0u; }
  unsigned long long pre = 0ull;   // survivors in lower lanes of the same warp and sub-tile: eight 5-bit fields
#pragma unroll
  for (int k = 0; k < COMPACT_ITEMS; ++k) {
    const unsigned b = __ballot_sync(IMC_FULL_MASK, (alive_mask >> k) & 1u);
    pre |= (unsigned long long)__popc(b & ((1u << lane) - 1)) << (5 * k);
    if (lane == 0) warp_cnt[k][wid] = __popc(b);
  }
  __syncthreads();
  if (alive_mask == 0u) return;
  long long out = block_offs[blockIdx.x];
#pragma unroll 1
  for (int k = 0; k < COMPACT_ITEMS; ++k) {
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) { const int cw = warp_cnt[k][w]; total += cw; before += w < wid ? cw : 0; }
    if ((alive_mask >> k) & 1u) {
      const long long i = i0 + (long long)k * COMPACT_THREADS, o = out + before + (int)((pre >> (5 * k)) & 31ull);
      dst.t[o] = src.t[i]; dst.x[o] = src.x[i]; dst.mu[o] = src.mu[i]; dst.E[o] = src.E[i]; dst.E0[o] = src.E0[i];
      dst.cx[o] = src.cx[i]; dst.ks[o] = src.ks[i]; dst.id[o] = src.id[i];
      if (geom == 2) { dst.y[o] = src.y[i]; dst.cy[o] = src.cy[i]; } else dst.origin[o] = src.origin[i];
    }
    out += total;
  }
}

// ======================================================================================
// Tally.tally
// ======================================================================================
// census radiation energy density: E / (dx [*dy] * scale) per surviving particle (imc_tally.jl:92, :106).
// The particle list is close to cell order (new particles are emitted cell by cell, compaction is stable), so the
// lanes of a warp mostly hold the same cell: runs of consecutive lanes with equal cells are summed with a segmented
// shuffle scan and the last lane of each run issues the one atomic (Float64 sums in ATOMIC mode, exact integer sums in
// FIXED mode — both order-free to the tolerance / exactly, like the per-lane atomics they replace).
// four consecutive elements of a particle field with one load (8-byte fields: two 16-byte loads); i is a multiple of 4 and
// the field arrays come from cudaMalloc, so the address is aligned
template <class T> struct alignas(sizeof(T) * 4 > 16 ? 16 : sizeof(T) * 4) Pack4 { T v[4]; };
template <class T> __device__ __forceinline__ Pack4<T> load4(const T* __restrict__ p, long long i) { return *reinterpret_cast<const Pack4<T>*>(p + i); }

// the ATOMIC / FIXED body of k_census_tally; V = double (float tallies) or long long (FIXED)
template <class P, class V>
__device__ __forceinline__ void census_groups(const MeshDev<P>& m, const Parts<P>& p, long long n, const TallyArgs& ta, Tally<P>& tal) {
  using N = Num<P>;
  using S = typename P::store_t;
  constexpr bool fixed = std::is_integral<V>::value;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  auto deposit = [&](int cell, V x) {
    if constexpr (fixed) {
      if (ta.use_smem) smem_add64(&tal.s_fx[cell], (unsigned long long)x);
      else atomicAdd(reinterpret_cast<unsigned long long*>(ta.g_fx) + cell, (unsigned long long)x);
    } else {
      if (ta.use_smem) atomicAdd(&tal.s_acc[cell], (typename AccType<P>::type)x);
      else atomicAdd(ta.g_acc + cell, x);
    }
  };
  const long long n4 = (n + 3) >> 2;
  const long long n4_round = (n4 + 31) & ~31ll;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n4_round; g += stride) {
    ThreadRuns<V> r;
    if (g < n4) {
      const long long i0 = g << 2;
      S e4[4], f4[4]; int cx4[4], cy4[4]; unsigned char k4[4];
      if (i0 + 3 < n) {
        const Pack4<S> pe = load4(p.E, i0); const Pack4<int> pc = load4(p.cx, i0); const Pack4<unsigned char> pk = load4(p.ks, i0);
#pragma unroll
        for (int j = 0; j < 4; ++j) { e4[j] = pe.v[j]; f4[j] = pe.v[j]; cx4[j] = pc.v[j]; k4[j] = pk.v[j]; cy4[j] = 0; }
        if (m.geom == 1) { const Pack4<S> pf = load4(p.E0, i0);
#pragma unroll
          for (int j = 0; j < 4; ++j) f4[j] = pf.v[j]; }
        else { const Pack4<int> py = load4(p.cy, i0);
#pragma unroll
          for (int j = 0; j < 4; ++j) cy4[j] = py.v[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const long long i = i0 + j;
          const bool in = i < n;
          e4[j] = in ? p.E[i] : P::pack((typename P::comp_t)-1); f4[j] = in ? (m.geom == 1 ? p.E0[i] : p.E[i]) : P::pack((typename P::comp_t)-1);
          cx4[j] = in ? p.cx[i] : 0; k4[j] = in ? p.ks[i] : (unsigned char)0; cy4[j] = (in && m.geom == 2) ? p.cy[i] : 0;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (P::unpack(f4[j]) == (typename P::comp_t)-1) continue;   // dead (slot 8: startenergy in 1-D, energy in 2-D)
        const N E(P::unpack(e4[j])), scale(m.scales[k4[j]]);
        int cj; N d;
        if (m.geom == 1) { cj = cx4[j]; d = E / (N::load(m.dx, cx4[j]) * scale); }
        else { cj = cx4[j] + m.nx * cy4[j]; d = E / ((N::load(m.dx, cx4[j]) * N::load(m.dy, cy4[j])) * scale); }
        V vj;
        if constexpr (fixed) vj = __double2ll_rn(d.d() * ta.fx_mul); else vj = d.d();
        r.push(cj, vj, deposit);
      }
    }
    bool want, want_f;
    warp_join_runs(lane, r, want, want_f);
    // (Taking turns with plain adds on a warp-private accumulator set — MATCH.ANY groups the lanes that name the same cell —
    // instead of these atomics, which are compare-and-swap loops for floats: measured slower, k_census_tally 0.67 -> 1.05 ms.)
    if (want) deposit(r.cell, r.v);
    if (want_f) deposit(r.cell_f, r.v_f);
  }
}

template <class P>
__global__ void __launch_bounds__(TRACK_THREADS) k_census_tally(MeshDev<P> m, Parts<P> p, long long n, TallyArgs ta) {
  using N = Num<P>;
  using S = typename P::store_t;
  extern __shared__ __align__(16) unsigned char smem[];
  Tally<P> tal(ta, smem);
  tal.zero();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  if (ta.mode == IMC_TALLY_EXACT) {   // one record per particle, in particle order (Q19)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      if (!particle_alive(p, i, m.geom)) { ta.rec_key[i] = 0x7fffffffu; ta.rec_val[i] = 0.0; continue; }
      N E = N::load(p.E, i), scale(m.scales[p.ks[i]]);
      int cx = p.cx[i];
      if (m.geom == 1) tal.add(cx, E / (N::load(m.dx, cx) * scale), i);
      else { int cy = p.cy[i]; tal.add((long long)cx + (long long)m.nx * cy, E / ((N::load(m.dx, cx) * N::load(m.dy, cy)) * scale), i); }
    }
    return;
  }
  // ATOMIC / FIXED.  A thread takes FOUR consecutive particles (vector loads of every field it needs); the runs of equal
  // cells among them, and from lane to lane, end in one atomic each (imc_warp_runs.cuh).  Float64 partial sums in ATOMIC
  // mode, exact integers in FIXED.
  if (ta.mode == IMC_TALLY_FIXED) census_groups<P, long long>(m, p, n, ta, tal);
  else census_groups<P, double>(m, p, n, ta, tal);
  tal.flush();
}

// ---- EXACT tally mode: per-cell reduction of the sorted deposit records in the reference's order ----------
// start[c] = first record of cell key c in the key-sorted record array (keys carry the wide flag in bit 31)
static __global__ void k_exact_bounds(const unsigned* __restrict__ keys, long long R, long long nacc, long long* __restrict__ start) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nacc) return;
  long long lo = 0, hi = R;  // lower_bound of c on (key & 0x7fffffff)
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if ((long long)(keys[mid] & 0x7fffffffu) < c) lo = mid + 1; else hi = mid;
  }
  start[c] = lo;
}
// Julia Base.sum (pairwise, 1024-element sequential leaves) of vals[first..last], evaluated by one thread
template <class P>
__device__ Num<P> jl_sum_serial(const double* __restrict__ vals, long long first, long long last) {
  using N = Num<P>;
  struct Frame { long long first, last; int state; N v1; };
  Frame st[48];
  int sp = 0;
  st[sp++] = {first, last, 0, N()};
  N ret;
  while (sp > 0) {
    Frame& f = st[sp - 1];
    if (f.state == 0) {
      if (f.first == f.last) { ret = N::from_d(vals[f.first]); --sp; }
      else if (f.last - f.first < 1024) {
        N v = N::from_d(vals[f.first]) + N::from_d(vals[f.first + 1]);
        for (long long i = f.first + 2; i <= f.last; ++i) v = v + N::from_d(vals[i]);
        ret = v; --sp;
      } else {
        long long mid = f.first + ((f.last - f.first) >> 1);
        f.state = 1;
        st[sp++] = {f.first, mid, 0, N()};
      }
    } else if (f.state == 1) {
      f.v1 = ret; f.state = 2;
      long long mid = f.first + ((f.last - f.first) >> 1);
      st[sp++] = {mid + 1, f.last, 0, N()};
    } else { ret = f.v1 + ret; --sp; }
  }
  return ret;
}
// ---- the same two reductions by a whole warp, for cells with many records (1-D decks at scale: 10^4-10^5 deposits per
// cell).  The additions keep the reference's order — they are a dependent chain — but the records are read 32 at a time,
// coalesced, and handed to the chain by shuffles; every lane carries the same accumulator.
constexpr int EXACT_WARP_MIN = 64;   // segments at least this long go to k_exact_reduce_warp
__device__ __forceinline__ bool exact_block_segment(long long len, int pairwise);   // ... or, pairwise and very long, to k_exact_reduce_block
// warp_seq_add / warp_seq_add_skip: imc_warp_reduce.cuh
// warp_jl_sum: imc_warp_reduce.cuh
template <class P>
__global__ void k_exact_reduce_warp(const unsigned* __restrict__ keys, const double* __restrict__ vals, const long long* __restrict__ start,
                                    long long nacc, int pairwise, int skip_stagnant, double* __restrict__ out) {
  using N = Num<P>;
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < nacc; c += nwarps) {
    const long long b = start[c], e = start[c + 1];
    if (e - b < EXACT_WARP_MIN || exact_block_segment(e - b, pairwise)) continue;
    N v;
    if (pairwise) v = warp_jl_sum<P>(vals, b, e - 1, lane);
    else if (skip_stagnant) v = warp_seq_add_skip<P>(N(), keys, vals, b, e - 1, lane);
    else v = warp_seq_add<P>(N(), true, keys, vals, b, e - 1, lane);   // v = zero(T); v += record ... (imc_transport.jl:101, :120)
    if (lane == 0) out[c] = v.d();
  }
}
// Julia's pairwise sum of a very long segment by a whole block: the recursion's leaves (< 1024 elements each, all on two
// adjacent depths: see k_jlsum_leaves) are independent, so the warps of the block sum them concurrently, each leaf in
// order (warp_seq_add); the tree is then folded bottom-up in shared memory exactly as the recursion combines it.
constexpr int EXACT_BLOCK_MIN = 8192;        // pairwise segments at least this long go to k_exact_reduce_block ...
constexpr int EXACT_BLOCK_MAX_DEPTH = 12;    // ... while their tree has at most 2^12 leaf slots (4 Mi records)
constexpr int EXACT_BLOCK_THREADS = 1024;
__device__ __forceinline__ int jl_sum_depth_dev(long long n) { int d = 0; while (n > 1024) { n = (n + 1) / 2; ++d; } return d; }
constexpr int EXACT_SEQBLOCK_MIN = 4096;     // sequential segments at least this long go to k_exact_reduce_seqblock
__device__ __forceinline__ bool exact_block_segment(long long len, int pairwise) {
  if (!pairwise) return len >= EXACT_SEQBLOCK_MIN;
  return len >= EXACT_BLOCK_MIN && jl_sum_depth_dev(len) <= EXACT_BLOCK_MAX_DEPTH;
}
template <class P>
__global__ void __launch_bounds__(EXACT_BLOCK_THREADS) k_exact_reduce_block(const double* __restrict__ vals, const long long* __restrict__ start,
                                                                           long long nacc, double* __restrict__ out) {
  using N = Num<P>;
  __shared__ typename P::comp_t part[1 << EXACT_BLOCK_MAX_DEPTH];
  __shared__ unsigned char valid[1 << EXACT_BLOCK_MAX_DEPTH];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (long long c = blockIdx.x; c < nacc; c += gridDim.x) {
    const long long b = start[c], len = start[c + 1] - b;
    if (!exact_block_segment(len, 1)) continue;
    const int depth = jl_sum_depth_dev(len);
    const int slots = 1 << depth;
    if (slots >= EXACT_BLOCK_THREADS / 8) {
      // many leaves (>= 128: 10^5 records and more in this cell): one THREAD per leaf.  Each leaf is summed left to right as the
      // recursion does, the leaves are independent chains, and 1024 of them run at once; a warp per leaf (below) walks one
      // chain with 32 lanes in lockstep, 40 cycles per record (census of the 10^8-particle Su-Olson deck: 3.5 ms vs 0.3 ms).
      // A thread's loads are 8 bytes at an 8 KB stride; the 32-byte sectors it pulls stay in L1 for its next three records.
      for (int slot = threadIdx.x; slot < slots; slot += blockDim.x) {
        long long first = 0, last = len - 1;
        int d = 0; bool mine = true;
        while (last - first >= 1024) {
          const long long mid = first + ((last - first) >> 1);
          if ((slot >> (depth - 1 - d)) & 1) first = mid + 1; else last = mid;
          ++d;
        }
        if (d < depth && (slot & ((1 << (depth - d)) - 1)) != 0) mine = false;
        if (mine) {
          N v = N::from_d(vals[b + first]);
          for (long long i = b + first + 1; i <= b + last; ++i) v = v + N::from_d(vals[i]);
          part[slot] = v.v;
        }
        valid[slot] = mine ? 1 : 0;
      }
    } else
    for (int slot = wid; slot < slots; slot += nw) {     // walk from the root to this slot's leaf (uniform in the warp)
      long long first = 0, last = len - 1;
      int d = 0; bool mine = true;
      while (last - first >= 1024) {
        const long long mid = first + ((last - first) >> 1);
        if ((slot >> (depth - 1 - d)) & 1) first = mid + 1; else last = mid;
        ++d;
      }
      if (d < depth && (slot & ((1 << (depth - d)) - 1)) != 0) mine = false;   // a shallower leaf belongs to its leftmost slot
      if (mine) {
        const double x0 = vals[b + first];
        const N v = first == last ? N::from_d(x0) : warp_seq_add<P>(N::from_d(x0), true, nullptr, vals, b + first + 1, b + last, lane);
        if (lane == 0) part[slot] = v.v;
      }
      if (lane == 0) valid[slot] = mine ? 1 : 0;
    }
    __syncthreads();
    for (int level = depth; level >= 1; --level) {         // k_jlsum_fold
      const int nodes = 1 << (level - 1), sh = depth - level;
      for (int i = threadIdx.x; i < nodes; i += blockDim.x) {
        const int ls = (2 * i) << sh, rs = (2 * i + 1) << sh;
        if (valid[rs]) part[ls] = (N(part[ls]) + N(part[rs])).v;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = (double)part[0];
    __syncthreads();
  }
}
// `v = zero(T); v += record ...` (imc_transport.jl:101, :120) over a very long segment by a whole block.  The additions
// are one dependent chain, but a Float16 / Float32 sum stops moving once it dwarfs the deposits (10^6 records per source
// cell of the Float16 Su-Olson deck, a few thousand of which change the sum), and "this record leaves v unchanged" can be
// tested for 1024 records at once: every thread adds ITS record to the running value; if no result differs from it the
// chain over the chunk leaves it unchanged too and the chunk is skipped.  Otherwise warp 0 walks the chunk from shared
// memory with warp_seq_add_skip (which jumps from one changing record to the next).  Same bits as the plain loop.
template <class P>
__global__ void __launch_bounds__(EXACT_BLOCK_THREADS) k_exact_reduce_seqblock(const unsigned* __restrict__ keys, const double* __restrict__ vals,
                                                                              const long long* __restrict__ start, long long nacc, double* __restrict__ out) {
  using N = Num<P>;
  __shared__ double s_x[EXACT_BLOCK_THREADS];
  __shared__ unsigned s_k[EXACT_BLOCK_THREADS];
  __shared__ int s_any[EXACT_BLOCK_THREADS / 32];
  __shared__ typename P::comp_t s_v;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (long long c = blockIdx.x; c < nacc; c += gridDim.x) {
    const long long b = start[c], len = start[c + 1] - b;
    if (!exact_block_segment(len, 0)) continue;
    N v;                                                                            // zero(T)
    for (long long base = 0; base < len; base += EXACT_BLOCK_THREADS) {
      const long long i = base + threadIdx.x;
      const bool have = i < len;
      double x = 0.0; unsigned k = 0u;
      if (have) { x = vals[b + i]; k = keys[b + i]; }
      const N t = (k & 0x80000000u) ? N::from_d(v.d() + x) : v + N::from_d(x);
      const bool same = (t.v == v.v && signbit(t.v) == signbit(v.v)) || (t.v != t.v && v.v != v.v);
      const unsigned m = __ballot_sync(IMC_FULL_MASK, have && !same);
      if (lane == 0) s_any[wid] = m != 0u;
      s_x[threadIdx.x] = x; s_k[threadIdx.x] = k;
      __syncthreads();
      if (wid == 0) {
        if (__ballot_sync(IMC_FULL_MASK, s_any[lane] != 0) != 0u) {
          const long long cnt = len - base < EXACT_BLOCK_THREADS ? len - base : EXACT_BLOCK_THREADS;
          v = warp_seq_add_skip<P>(v, s_k, s_x, 0, cnt - 1, lane);
        }
        if (lane == 0) s_v = v.v;
      }
      __syncthreads();
      v = N(s_v);
    }
    if (threadIdx.x == 0) out[c] = v.d();
  }
}
// one thread per tally cell: PAIRWISE = FALSE -> `+=` in record order (imc_transport.jl:101,120); TRUE -> sum(vector) (:202)
template <class P>
__global__ void k_exact_reduce(const unsigned* __restrict__ keys, const double* __restrict__ vals, const long long* __restrict__ start,
                               long long nacc, int pairwise, double* __restrict__ out) {
  using N = Num<P>;
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nacc) return;
  long long b = start[c], e = start[c + 1];
  if (e - b >= EXACT_WARP_MIN) return;   // long segments: k_exact_reduce_warp
  N v;
  if (e > b) {
    if (pairwise) v = jl_sum_serial<P>(vals, b, e - 1);
    else for (long long i = b; i < e; ++i) {
      if (keys[i] & 0x80000000u) v = N::from_d(v.d() + vals[i]);   // Float64 deposit added to a T accumulator (MC_RW)
      else v = v + N::from_d(vals[i]);
    }
  }
  out[c] = v.d();
}
// lost energy in particle order (imc_transport.jl:141 / :200).  The per-particle slots (NaN = no loss) are first
// compacted in order (k_lost_flags, scan, k_lost_gather) — escapes are few next to the population — and one thread
// then adds the compacted values exactly as the reference's loop does.
static __global__ void k_lost_flags(const double* __restrict__ lost_val, long long n, int* __restrict__ flag) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const double e = lost_val[i]; flag[i] = e == e ? 1 : 0; }
}
static __global__ void k_lost_gather(const double* __restrict__ lost_val, const unsigned char* __restrict__ ks, const long long* __restrict__ offs,
                                     long long n, double* __restrict__ cval, unsigned char* __restrict__ cks) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double e = lost_val[i];
  if (e == e) { cval[offs[i]] = e; cks[offs[i]] = ks[i]; }
}
template <class P>
__global__ void k_exact_lost(const double* __restrict__ lost_val, const unsigned char* __restrict__ ks, long long n, MeshDev<P> m,
                             int pairwise, double* __restrict__ scratch, double* __restrict__ lost_io) {
  using N = Num<P>;
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double lost = *lost_io; bool wide = false;
  if (!pairwise) {
    for (long long i = 0; i < n; ++i) {
      double e = lost_val[i];
      if (e != e) continue;
      N scale(m.scales[ks[i]]);
      N en = N::from_d(e);
      if (en.d() != e) { lost = lost + e / scale.d(); wide = true; }                 // Float64 energy (MC_RW): lostenergy turns Float64
      else if (wide) lost = lost + (en / scale).d();
      else lost = (N::from_d(lost) + en / scale).d();
    }
  } else {
    for (int k = 0; k < m.ns; ++k) {
      long long cnt = 0;
      for (long long i = 0; i < n; ++i) { double e = lost_val[i]; if (e == e && ks[i] == k) scratch[cnt++] = N::from_d(e).d(); }
      N sum = cnt ? jl_sum_serial<P>(scratch, 0, cnt - 1) : N();
      lost = (N::from_d(lost) + sum / N(m.scales[k])).d();
    }
  }
  *lost_io = lost;
}

// reduce buffer -> energydep (T).  kind: 0 Float64 accumulators, 1 fixed-point int64
// (nx, sx, sy): accumulator of cell (xi, yi) of plane k sits at k*nc + xi*sx + yi*sy (MeshDev::tsx / tsy); nx = 0: linear
template <class P>
__global__ void k_acc_to_field(const double* g_acc, int kind, double fx_mul, long long n, typename P::store_t* out,
                               long long nc, int nx, int sx, int sy) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long j = i;
  if (nx > 0) { const long long k = i / nc, c = i - k * nc; j = k * nc + (c % nx) * (long long)sx + (c / nx) * (long long)sy; }
  double v = kind == 1 ? (double)reinterpret_cast<const long long*>(g_acc)[j] / fx_mul : g_acc[j];
  Num<P>::from_d(v).store(out, i);
}

template <class P>
struct TallyScratch {
  typename P::store_t* q_dep;   // [Nc x Ns]  (energydep * vol) / scale
  typename P::store_t* q_tot;   // [Nc]       matenergydens + radenergydens
  typename P::store_t* q_rad;   // [Nc]       radenergydens * vol (energycheck)
};

template <class P>
__global__ void k_tally_finish(MeshDev<P> m, TallyScratch<P> s, typename P::comp_t dt_, int t_is_zero, int linearized, int temp_wide) {
  using N = Num<P>;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.nc) return;
  N dt(dt_);
  double t = m.temp[i];
  N mat = N::load(m.matenergydens, i);
  if (t_is_zero) {                                                                 // :29-32 (Q11)
    double vals[10] = {N::load(m.fleck, i).d(), N::load(m.sa, i).d(), (double)m.a, (double)m.c, t, t, t, t, dt.d(), (double)m.ds};
    double sc1 = 1.0; int idx;
    sorter_dev<P, 10>(vals, &sc1, 1, &mat, &idx);
  }
  int xi = (int)(m.geom == 1 ? i : i % m.nx), yi = (int)(m.geom == 1 ? 0 : i / m.nx);
  N vol = m.geom == 1 ? N::load(m.dx, xi) : N::load(m.dx, xi) * N::load(m.dy, yi);
  N inc;
  for (int k = 0; k < m.ns; ++k) {                                                 // :47-57
    N dep = N::load(m.energydep, i + m.nc * k), em = N::load(m.emittedenergy, i + m.nc * k), sc(m.scales[k]);
    inc = inc + (dep - em) / sc;
    ((dep * vol) / sc).store(s.q_dep, i + m.nc * k);
  }
  inc.store(m.nrg_inc, i);
  mat = mat + inc;                                                                 // :68
  mat.store(m.matenergydens, i);
  if (linearized) t = MathDet::pow64(mat.d(), 0.25);                               // :72 (Q12)
  else t = temp_wide ? t + (inc / N::load(m.bee, i)).d() : (N::from_d(t) + inc / N::load(m.bee, i)).d();  // :74
  m.temp[i] = t;
  (mat + N::load(m.radenergydens, i)).store(s.q_tot, i);
}

template <class P>
__global__ void k_rad_energy(MeshDev<P> m, typename P::store_t* q_rad) {          // imc_energycheck.jl:24-29
  using N = Num<P>;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.nc) return;
  N r = N::load(m.radenergydens, i);
  if (m.geom == 1) (r * N::load(m.dx, i)).store(q_rad, i);
  else ((r * N::load(m.dx, i % m.nx)) * N::load(m.dy, i / m.nx)).store(q_rad, i);
}

// maximum(temp): NaN-propagating like Julia's maximum
static __global__ void k_max_f64(const double* __restrict__ v, long long n, double* out, int* has_nan) {
  double mx = -INFINITY; int nan = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double x = v[i];
    if (x != x) nan = 1; else if (x > mx) mx = x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { double y = __shfl_xor_sync(IMC_FULL_MASK, mx, o); if (y > mx) mx = y; nan |= __shfl_xor_sync(IMC_FULL_MASK, nan, o); }
  if ((threadIdx.x & 31) == 0) {
    if (nan) atomicOr(has_nan, 1);
    // atomic max on doubles through the ordered integer image
    unsigned long long* addr = reinterpret_cast<unsigned long long*>(out);
    unsigned long long old = *addr, assumed;
    do {
      assumed = old;
      if (__longlong_as_double((long long)assumed) >= mx) break;
      old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(mx));
    } while (assumed != old);
  }
}

// Scalars a stage hands back to the host live in several device arrays (reduction slots, the reduce buffer's tail, flags).
// One launch gathers up to GATHER_MAX of them as 8-byte words (values of T and ints converted to Float64, 8-byte values
// copied bit for bit) so that the host needs ONE copy into pinned memory and one synchronisation per stage — each separate
// cudaMemcpyAsync into pageable host memory is a synchronisation of its own.
constexpr int GATHER_MAX = 32;
enum { GK_T = 0, GK_RAW8 = 1, GK_I32 = 2 };
struct GatherList { const void* p[GATHER_MAX]; int kind[GATHER_MAX]; int count; };
template <class P>
__global__ void k_gather_scalars(GatherList g, double* __restrict__ out) {
  const int i = threadIdx.x;
  if (i >= g.count) return;
  if (g.kind[i] == GK_T) out[i] = (double)*static_cast<const typename P::comp_t*>(g.p[i]);
  else if (g.kind[i] == GK_I32) out[i] = (double)*static_cast<const int*>(g.p[i]);
  else out[i] = *static_cast<const double*>(g.p[i]);
}

// fixed-point slot: value * ratio (a power of two), when the scale of the slot changes between two calls
static __global__ void k_rescale_fixed(long long* slot, double ratio) { *slot = __double2ll_rn((double)*slot * ratio); }

// max particle energy (for the fixed-point scale)
template <class P>
__global__ void k_max_energy(Parts<P> p, long long n, double* out) {
  double mx = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double e = (double)P::unpack(p.E[i]);
    if (e > mx) mx = e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { double y = __shfl_xor_sync(IMC_FULL_MASK, mx, o); if (y > mx) mx = y; }
  if ((threadIdx.x & 31) == 0 && mx > 0) {
    unsigned long long* addr = reinterpret_cast<unsigned long long*>(out);
    unsigned long long old = *addr, assumed;
    do {
      assumed = old;
      if (__longlong_as_double((long long)assumed) >= mx) break;
      old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(mx));
    } while (assumed != old);
  }
}

template <class P>
__global__ void k_to_f64(const typename P::store_t* __restrict__ src, long long n, double* __restrict__ dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)P::unpack(src[i]);
}
template <class P>
__global__ void k_from_f64(const double* __restrict__ src, long long n, typename P::store_t* __restrict__ dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = P::pack(P::from_d(src[i]));
}
template <class P>
__global__ void k_round_f64(const double* __restrict__ src, long long n, double* __restrict__ dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)P::from_d(src[i]);
}

// particle SoA <-> reference slot layout (imc_get_particles / imc_set_particles)
template <class P>
__global__ void k_export_particles(MeshDev<P> m, Parts<P> p, long long n, double* slots, unsigned long long* ids) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  using N = Num<P>;
  double scale = (double)m.scales[p.ks[i]];
  if (m.geom == 1) {
    double* s = slots + i * 9;
    s[0] = (double)(p.origin[i] + 1); s[1] = N::load(p.t, i).d(); s[2] = (double)(p.cx[i] + 1); s[3] = N::load(p.x, i).d();
    s[4] = N::load(p.mu, i).d(); s[5] = 1.0; s[6] = N::load(p.E, i).d(); s[7] = N::load(p.E0, i).d(); s[8] = scale;
  } else {
    double* s = slots + i * 10;
    s[0] = N::load(p.t, i).d(); s[1] = (double)(p.cx[i] + 1); s[2] = (double)(p.cy[i] + 1); s[3] = N::load(p.x, i).d();
    s[4] = N::load(p.y, i).d(); s[5] = N::load(p.mu, i).d(); s[6] = 1.0; s[7] = N::load(p.E, i).d(); s[8] = N::load(p.E0, i).d(); s[9] = scale;
  }
  if (ids) ids[i] = p.id[i];
}
template <class P>
__global__ void k_import_particles(MeshDev<P> m, Parts<P> p, long long n, const double* slots, const unsigned long long* ids, int* bad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  using N = Num<P>;
  const double* s = slots + i * (m.geom == 1 ? 9 : 10);
  double scale = s[m.geom == 1 ? 8 : 9];
  int ks = -1;
  for (int k = 0; k < m.ns; ++k) if ((double)m.scales[k] == scale) { ks = k; break; }   // findfirst(isequal(scale), energyscales)
  if (ks < 0) { atomicOr(bad, 1); ks = 0; }
  p.ks[i] = (unsigned char)ks;
  p.id[i] = ids ? ids[i] : (unsigned long long)i;
  if (m.geom == 1) {
    p.origin[i] = (int)s[0] - 1; N::from_d(s[1]).store(p.t, i); p.cx[i] = (int)s[2] - 1; N::from_d(s[3]).store(p.x, i);
    N::from_d(s[4]).store(p.mu, i); N::from_d(s[6]).store(p.E, i); N::from_d(s[7]).store(p.E0, i);
    if (p.cx[i] < 0 || p.cx[i] >= m.nx) atomicOr(bad, 2);
  } else {
    N::from_d(s[0]).store(p.t, i); p.cx[i] = (int)s[1] - 1; p.cy[i] = (int)s[2] - 1; N::from_d(s[3]).store(p.x, i);
    N::from_d(s[4]).store(p.y, i); N::from_d(s[5]).store(p.mu, i); N::from_d(s[7]).store(p.E, i); N::from_d(s[8]).store(p.E0, i);
    if (p.cx[i] < 0 || p.cx[i] >= m.nx || p.cy[i] < 0 || p.cy[i] >= m.ny) atomicOr(bad, 2);
  }
}

}  // namespace imc
