// imc_engine_impl.cuh — EngineT<P>: host orchestration of the transport-step kernels for one GPU.
// Owns all device memory (mesh fields, particle SoA double buffer, reduce buffer, scratch).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "imc_engine.h"
#include "imc_kernels.cuh"
#include "imc_exact.h"

namespace imc {

#define IMC_CK(call)                                                                               \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char b_[512];                                                                                \
      snprintf(b_, sizeof b_, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
      err = b_;                                                                                    \
      return e_ == cudaErrorMemoryAllocation ? IMC_ERR_NOMEM : IMC_ERR_CUDA;                        \
    }                                                                                              \
  } while (0)
#define IMC_RC(call) do { int rc_ = (call); if (rc_) return rc_; } while (0)

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count, bool zero = true) {
    release();
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e != cudaSuccess) { p = nullptr; return e; }
    n = count;
    return zero ? cudaMemset(p, 0, count * sizeof(T)) : cudaSuccess;
  }
  // scratch that follows the population (records, flags, scans): grow by at least a quarter, so that a population creeping
  // upward does not pay a cudaFree + cudaMalloc (and its device synchronisation) every time step
  cudaError_t ensure(size_t count) { return count <= n ? cudaSuccess : alloc(count > n + n / 4 ? count : n + n / 4, false); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DBuf() { release(); }
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
};

inline unsigned grid_for(long long n, int threads) { return (unsigned)std::max<long long>(1, (n + threads - 1) / threads); }

template <class P>
struct PartBufs {
  using S = typename P::store_t;
  DBuf<S> t, x, y, mu, E, E0;
  DBuf<int> cx, cy, origin;
  DBuf<unsigned char> ks;
  DBuf<unsigned long long> id;
  cudaError_t alloc(size_t cap, int geom) {
    cudaError_t e;
#define A_(b) if ((e = b.alloc(cap, false)) != cudaSuccess) return e
    A_(t); A_(x); A_(mu); A_(E); A_(E0); A_(cx); A_(ks); A_(id);
    if (geom == 2) { A_(y); A_(cy); } else { A_(origin); }
#undef A_
    return cudaSuccess;
  }
  Parts<P> view() { Parts<P> v; v.t = t.p; v.x = x.p; v.y = y.p; v.mu = mu.p; v.E = E.p; v.E0 = E0.p; v.cx = cx.p; v.cy = cy.p; v.origin = origin.p; v.ks = ks.p; v.id = id.p; return v; }
};

template <class P>
struct EngineT : EngineBase {
  using S = typename P::store_t;
  using Cc = typename P::comp_t;
  using N = Num<P>;
  imc_config cfg;
  int geom, nx, ny, ns, sm_count = 148;
  long long nc;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool have_mesh = false, temp_wide = false, have_rw = false;
  // mesh
  DBuf<S> dx, dy, wx, wy, sa, ss, fleck, beta, bee, sa_c, sa_p, ss_c, ss_p, sigma_static, radsource;
  DBuf<S> matenergydens, radenergydens, nrg_inc, energydep, emittedenergy, tsurf[4];
  DBuf<double> temp;
  DBuf<CellProp1<P>> cp1;
  DBuf<CellProp2<P>> cp2;
  DBuf<AxisProp<P>> ax_inv, ax_d;   // x entries, then y entries
  MeshDev<P> m;
  // particles (double buffer for the stable compaction)
  PartBufs<P> pb[2];
  int cur = 0;
  long long n_part = 0, cap = 0;
  // reduce buffer: [energydep Nc*Ns | radenergydens Nc | scalars]
  DBuf<double> red;
  long long red_n = 0;
  bool red_fixed = false;
  bool y_fastest = false, dep_perm = false;   // layout of the CellProp2 table / of the deposit accumulators of the last transport (MeshDev::csx ...)
  double fx_mul_dep = 1, fx_mul_rad = 1, fx_mul_lost = 1;
  // sourcing scratch
  SrcLayout L;
  DBuf<S> src_e, src_q, src_nrg, src_qem;
  DBuf<signed char> src_ks;
  DBuf<int> src_cnt;
  DBuf<long long> src_offs, src_block_entry, scan_tiles, scan_tiles2, scan_total;
  DBuf<SrcScalars> src_sc;
  DBuf<Cc> sums;  // device slots for jl_sum results
  // jl_sum scratch
  DBuf<Cc> jl_part;
  DBuf<unsigned char> jl_valid;
  // tally scratch
  DBuf<S> q_dep, q_tot, q_rad;
  DBuf<double> d_max;
  DBuf<int> d_flag;
  DBuf<unsigned long long> over_flag, timeline;
  DBuf<long long> blk_cnt;
  // EXACT tally mode: deposit records
  DBuf<int> rec_cnt;
  DBuf<long long> rec_off, rec_start;
  DBuf<unsigned> rec_key[2];
  DBuf<double> rec_val[2], lost_val, lost_scratch, lost_cval;
  DBuf<unsigned char> lost_cks;
  DBuf<unsigned char> sort_temp;
  int last_mode = IMC_TALLY_ATOMIC;
  // event-based schedule
  DBuf<unsigned> ev_list[2], ev_extra;
  DBuf<int> ev_nseg;
  DBuf<unsigned long long> ev_count;
  double rate_event = 0;
  int ev_launches = 0;
  // outcomes
  DBuf<signed char> out_event;
  DBuf<int> out_nseg;
  long long out_n = 0;
  // tapes
  DBuf<double> tt_uni, tt_exp, st_uni;
  int tt_nuni = 0, tt_nexp = 0, st_nuni = 0;
  long long tt_slots = 0, st_slots = 0;
  // random-walk tables
  DBuf<S> rw_a, rw_pt;
  std::vector<double> h_rw_a, h_rw_pr, h_rw_pt;
  // host mirrors of the scalar state
  double totalenergy = 0, totalenergydep = 0, radenergyold = 0;
  uint64_t iterations = 0;
  long long n_transport_calls = 0;
  // replicated quantities behind the EXACT-or-not decision of AUTO (identical on every rank of a multi-GPU run): the segments
  // all ranks tracked in the previous step (read from the all-reduced buffer by tally_finish) and the global population
  // after sourcing
  double last_global_segments = 0; long long n_global_after_source = 0;
  double rate_static = 0, rate_refill = 0;  // segments per ms of each schedule, last measured
  static int refill_min_env() { const char* e = getenv("IMC_REFILL_MIN"); int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > 32 ? 32 : v); }
  int64_t n_launch = 0;  // kernels launched by this engine (bench.py reports it as gpu_launches)

  explicit EngineT(const imc_config& c) : cfg(c) {
    geom = c.geometry; nx = c.nx; ny = geom == 2 ? c.ny : 1; ns = c.n_scales; nc = (long long)nx * ny;
    L.geom = geom; L.nx = nx; L.ny = ny; L.nc = nc;
  }
  ~EngineT() override {
    if (hpin) cudaFreeHost(hpin);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }

  int init() override {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
      err = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); the transport step has no CPU fallback";
      return IMC_ERR_CUDA;
    }
    if (cfg.device < 0 || cfg.device >= ndev) { err = "bad device ordinal"; return IMC_ERR_ARG; }
    if (nc >= (1ll << 30)) { err = "mesh has 2^30 or more cells (cell and source-entry indices are 32-bit)"; return IMC_ERR_ARG; }
    if (nc * ns >= (1ll << 31)) { err = "cells x energy scales reaches 2^31 (tally accumulator indices are 32-bit)"; return IMC_ERR_ARG; }
    IMC_CK(cudaSetDevice(cfg.device));
    cudaDeviceProp prop;
    IMC_CK(cudaGetDeviceProperties(&prop, cfg.device));
    sm_count = prop.multiProcessorCount;
    IMC_CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    IMC_CK(cudaEventCreate(&ev0));
    IMC_CK(cudaEventCreate(&ev1));
    return IMC_OK;
  }
  int use_device() { IMC_CK(cudaSetDevice(cfg.device)); return IMC_OK; }
  // one gather launch + one copy into pinned memory + one synchronisation per stage (k_gather_scalars)
  double* hpin = nullptr;
  DBuf<double> dgather;
  GatherList gl{};
  void gl_reset() { gl.count = 0; }
  int gl_add(const void* p, int kind) {
    if (gl.count >= GATHER_MAX) return -1;   // callers request at most 6 + IMC_MAX_SCALES values
    gl.p[gl.count] = p; gl.kind[gl.count] = kind; return gl.count++;
  }
  int gl_read() {   // results in hpin[0 .. count)
    if (!hpin) { IMC_CK(cudaMallocHost((void**)&hpin, GATHER_MAX * sizeof(double))); IMC_CK(dgather.alloc(GATHER_MAX)); }
    k_gather_scalars<P><<<1, GATHER_MAX, 0, stream>>>(gl, dgather.p); ++n_launch;
    IMC_CK(cudaMemcpyAsync(hpin, dgather.p, gl.count * sizeof(double), cudaMemcpyDeviceToHost, stream));
    IMC_CK(cudaStreamSynchronize(stream));
    return IMC_OK;
  }
  static long long raw_ll(double v) { long long r; memcpy(&r, &v, 8); return r; }

  // ---- helpers -----------------------------------------------------------------------------
  int upload(DBuf<S>& b, const double* src, size_t n) {
    std::vector<S> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = P::pack(src ? P::from_d(src[i]) : (Cc)0);
    IMC_CK(b.alloc(n, false));
    IMC_CK(cudaMemcpyAsync(b.p, h.data(), n * sizeof(S), cudaMemcpyHostToDevice, stream));
    IMC_CK(cudaStreamSynchronize(stream));
    return IMC_OK;
  }
  DBuf<double> stage;  // device staging for Float64 <-> T conversion at the ABI boundary
  int download(const S* src, size_t n, double* dst) {
    IMC_CK(stage.ensure(n));
    k_to_f64<P><<<grid_for((long long)n, 256), 256, 0, stream>>>(src, (long long)n, stage.p); ++n_launch;
    IMC_CK(cudaMemcpyAsync(dst, stage.p, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    IMC_CK(cudaStreamSynchronize(stream));
    return IMC_OK;
  }
  int upload_into(S* dst, const double* src, size_t n) {
    IMC_CK(stage.ensure(n));
    IMC_CK(cudaMemcpyAsync(stage.p, src, n * sizeof(double), cudaMemcpyHostToDevice, stream));
    k_from_f64<P><<<grid_for((long long)n, 256), 256, 0, stream>>>(stage.p, (long long)n, dst); ++n_launch;
    IMC_CK(cudaGetLastError());
    return IMC_OK;
  }
  // Julia sums of q[0..n) into device slots: requests are queued and run as one launch pair per batch (jl_flush)
  JlSumBatch<P> jl_batch{};
  int jl_sum(const S* q, long long n, Cc* out) {
    if (jl_batch.count == JLSUM_BATCH) IMC_RC(jl_flush());
    const int k = jl_batch.count++;
    jl_batch.q[k] = q; jl_batch.n[k] = n; jl_batch.depth[k] = jl_sum_depth(n); jl_batch.out[k] = out;
    return IMC_OK;
  }
  int jl_flush() {
    if (jl_batch.count == 0) return IMC_OK;
    int dmax = 0;
    for (int k = 0; k < jl_batch.count; ++k) dmax = std::max(dmax, jl_batch.depth[k]);
    const size_t slots = (size_t)1 << dmax;
    jl_batch.slots_max = (long long)slots;
    IMC_CK(jl_part.ensure(slots * JLSUM_BATCH));
    IMC_CK(jl_valid.ensure(slots * JLSUM_BATCH));
    dim3 grid(grid_for((long long)slots, 128), (unsigned)jl_batch.count);
    k_jlsum_leaves_multi<P><<<grid, 128, 0, stream>>>(jl_batch, jl_part.p, jl_valid.p); ++n_launch;
    k_jlsum_fold_multi<P><<<(unsigned)jl_batch.count, 1024, 0, stream>>>(jl_batch, jl_part.p, jl_valid.p); ++n_launch;
    jl_batch.count = 0;
    IMC_CK(cudaGetLastError());
    return IMC_OK;
  }
  // exclusive scan int32 -> int64, total to scan_total.p[0]
  int scan_counts(const int* in, long long* out, long long n) {
    long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    IMC_CK(scan_tiles.ensure((size_t)tiles));
    IMC_CK(scan_total.ensure(1));
    k_scan_tiles<int><<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(in, out, n, scan_tiles.p); ++n_launch;
    k_scan_small<<<1, 1024, 0, stream>>>(scan_tiles.p, tiles, scan_total.p); ++n_launch;
    k_scan_add<<<grid_for(n, 256), 256, 0, stream>>>(out, n, scan_tiles.p); ++n_launch;
    IMC_CK(cudaGetLastError());
    return IMC_OK;
  }
  int ensure_capacity(long long need) {
    if (need <= cap) return IMC_OK;
    long long ncap = std::max<long long>(need, cap + cap / 2);
    ncap = std::max<long long>(ncap, 1024);
    PartBufs<P> nb[2];
    IMC_CK(nb[0].alloc((size_t)ncap, geom));
    IMC_CK(nb[1].alloc((size_t)ncap, geom));
    if (n_part > 0) {
      PartBufs<P>& o = pb[cur];
#define CP_(f, T_) IMC_CK(cudaMemcpyAsync(nb[0].f.p, o.f.p, (size_t)n_part * sizeof(T_), cudaMemcpyDeviceToDevice, stream))
      CP_(t, S); CP_(x, S); CP_(mu, S); CP_(E, S); CP_(E0, S); CP_(cx, int); CP_(ks, unsigned char); CP_(id, unsigned long long);
      if (geom == 2) { CP_(y, S); CP_(cy, int); } else { CP_(origin, int); }
#undef CP_
      IMC_CK(cudaStreamSynchronize(stream));
    }
    for (int b = 0; b < 2; ++b) {
#define MV_(f) std::swap(pb[b].f.p, nb[b].f.p); std::swap(pb[b].f.n, nb[b].f.n)
      MV_(t); MV_(x); MV_(y); MV_(mu); MV_(E); MV_(E0); MV_(cx); MV_(cy); MV_(origin); MV_(ks); MV_(id);
#undef MV_
    }
    if (cur == 1) {  // live data was copied into buffer 0
      cur = 0;
    }
    cap = ncap;
    return IMC_OK;
  }
  RngArgs rng_args(int64_t step, bool source) {
    RngArgs r;
    r.tape = cfg.rng_mode == IMC_RNG_TAPE;
    r.seed = (unsigned long long)cfg.seed;
    Philox::round_keys(r.seed, r.rk);
    r.step = (unsigned int)step;
    if (source) { r.uni = st_uni.p; r.ex = nullptr; r.n_uni = st_nuni; r.n_exp = 0; r.stride = st_slots; }
    else { r.uni = tt_uni.p; r.ex = tt_exp.p; r.n_uni = tt_nuni; r.n_exp = tt_nexp; r.stride = tt_slots; }
    return r;
  }
  long long rb_dep0() const { return 0; }
  long long rb_rad0() const { return nc * ns; }
  long long rb_sc0() const { return nc * ns + nc; }

  // ---- set_mesh ----------------------------------------------------------------------------
  int set_mesh(const double* dx_, const double* dy_, const double* sac, const double* sap, const double* ssc,
               const double* ssp, const double* sstat, const double* bee_, const double* rad, const double* temp_,
               const double* tsb, const double* tst, const double* tsl, const double* tsr) override {
    IMC_RC(use_device());
    if (!dx_ || !sac || !sap || !ssc || !ssp || !bee_ || !rad || !temp_ || !tsl || !tsr || (geom == 2 && (!dy_ || !tsb || !tst))) {
      err = "set_mesh: null array"; return IMC_ERR_ARG;
    }
    IMC_RC(upload(dx, dx_, nx));
    if (geom == 2) IMC_RC(upload(dy, dy_, ny)); else { double one = 1.0; IMC_RC(upload(dy, &one, 1)); }
    IMC_CK(wx.alloc(nx)); IMC_CK(wy.alloc(ny)); IMC_CK(ax_inv.alloc(nx + ny)); IMC_CK(ax_d.alloc(nx + ny));
    IMC_RC(upload(sa_c, sac, nc)); IMC_RC(upload(sa_p, sap, nc)); IMC_RC(upload(ss_c, ssc, nc)); IMC_RC(upload(ss_p, ssp, nc));
    IMC_RC(upload(sa, sac, nc)); IMC_RC(upload(ss, ssc, nc));
    IMC_RC(upload(sigma_static, sstat, nc));
    IMC_RC(upload(bee, bee_, nc)); IMC_RC(upload(radsource, rad, nc));
    {
      std::vector<double> h(nc);
      for (long long i = 0; i < nc; ++i) h[i] = (double)P::from_d(temp_[i]);
      IMC_CK(temp.alloc(nc, false));
      IMC_CK(cudaMemcpy(temp.p, h.data(), nc * sizeof(double), cudaMemcpyHostToDevice));
    }
    temp_wide = false;
    if (geom == 1) { IMC_RC(upload(tsurf[2], tsl, 1)); IMC_RC(upload(tsurf[3], tsr, 1)); IMC_CK(tsurf[0].alloc(1)); IMC_CK(tsurf[1].alloc(1)); }
    else { IMC_RC(upload(tsurf[0], tsb, nx)); IMC_RC(upload(tsurf[1], tst, nx)); IMC_RC(upload(tsurf[2], tsl, ny)); IMC_RC(upload(tsurf[3], tsr, ny)); }
    IMC_CK(fleck.alloc(nc)); IMC_CK(beta.alloc(nc));
    IMC_CK(matenergydens.alloc(nc)); IMC_CK(radenergydens.alloc(nc)); IMC_CK(nrg_inc.alloc(nc));
    IMC_CK(energydep.alloc(nc * ns)); IMC_CK(emittedenergy.alloc(nc * ns));
    if (geom == 1) IMC_CK(cp1.alloc(nc)); else IMC_CK(cp2.alloc(nc));
    red_n = nc * ns + nc + RB_NSCALARS;
    IMC_CK(red.alloc(red_n));
    // sourcing / tally scratch
    long long M = L.total();
    IMC_CK(src_e.alloc(M)); IMC_CK(src_q.alloc(M)); IMC_CK(src_nrg.alloc(M)); IMC_CK(src_ks.alloc(M)); IMC_CK(src_cnt.alloc(M));
    IMC_CK(src_offs.alloc(M + 1)); IMC_CK(src_qem.alloc(nc * ns)); IMC_CK(src_sc.alloc(1)); IMC_CK(sums.alloc(16 + IMC_MAX_SCALES));
    IMC_CK(q_dep.alloc(nc * ns)); IMC_CK(q_tot.alloc(nc)); IMC_CK(q_rad.alloc(nc));
    IMC_CK(d_max.alloc(2)); IMC_CK(d_flag.alloc(2)); IMC_CK(over_flag.alloc(2));
    // device view
    m.geom = geom; m.nx = nx; m.ny = ny; m.ns = ns; m.nc = nc;
    m.dx = dx.p; m.dy = dy.p; m.wx = wx.p; m.wy = wy.p;
    m.sa = sa.p; m.ss = ss.p; m.fleck = fleck.p; m.beta = beta.p; m.bee = bee.p; m.sa_c = sa_c.p; m.sa_p = sa_p.p;
    m.ss_c = ss_c.p; m.ss_p = ss_p.p; m.sigma_static = sigma_static.p; m.radsource = radsource.p; m.temp = temp.p;
    m.matenergydens = matenergydens.p; m.radenergydens = radenergydens.p; m.nrg_inc = nrg_inc.p;
    m.energydep = energydep.p; m.emittedenergy = emittedenergy.p;
    for (int k = 0; k < 4; ++k) m.tsurf[k] = tsurf[k].p;
    m.cp1 = cp1.p; m.cp2 = cp2.p; m.ax_inv = ax_inv.p; m.ax_d = ax_d.p;
    for (int k = 0; k < IMC_MAX_SCALES; ++k) { m.scales[k] = k < ns ? P::from_d(cfg.energyscales[k]) : (Cc)1; m.scales_d[k] = (double)m.scales[k]; }
    m.ds = P::from_d(cfg.distancescale); m.c = P::from_d(cfg.phys_c); m.a = P::from_d(cfg.phys_a); m.alpha = P::from_d(cfg.alpha);
    for (int k = 0; k < 4; ++k) m.bc[k] = cfg.bc[k];
    m.ds_is_one = (double)m.ds == 1.0; m.c_is_one = (double)m.c == 1.0;
    m.n_tdiv = 0; m.tdiv[0] = m.tdiv[1] = (Cc)1;
    if (!m.ds_is_one) m.tdiv[m.n_tdiv++] = m.ds;
    if (!m.c_is_one) m.tdiv[m.n_tdiv++] = m.c;
    for (int r = 0; r < 2; ++r) {   // RN(1/divisor) for the cached-reciprocal division (imc_fastdiv.cuh); 0 outside its range
      const double d = std::fabs((double)m.tdiv[r]);
      m.tdiv_r[r] = (P::id != 2 && d >= 0x1p-40 && d < 0x1p40) ? (Cc)(1.0f / (float)m.tdiv[r]) : (Cc)0;
    }
    // internal layout of the per-cell tracking tables (MeshDev::csx ...): the more frequently crossed axis fastest
    {
      double lx = 0, ly = 0;
      for (int i = 0; i < nx; ++i) lx += std::fabs(dx_[i]);
      if (geom == 2) for (int j = 0; j < ny; ++j) ly += std::fabs(dy_[j]);
      static const int force = getenv("IMC_CELL_ORDER") ? atoi(getenv("IMC_CELL_ORDER")) : 0;   // 1: x fastest, 2: y fastest (experiments)
      y_fastest = geom == 2 && (force ? force == 2 : (double)ny * lx > (double)nx * ly);   // ny / ly > nx / lx: y faces are crossed more often
      m.csx = y_fastest ? ny : 1; m.csy = y_fastest ? 1 : nx;
      m.tsx = 1; m.tsy = nx;
    }
    k_widths<P><<<grid_for(std::max(nx, ny), 256), 256, 0, stream>>>(m); ++n_launch;
    IMC_CK(cudaGetLastError());
    IMC_CK(cudaStreamSynchronize(stream));
    totalenergy = totalenergydep = radenergyold = 0;
    echeck_cached = false;
    have_mesh = true;
    return IMC_OK;
  }

  // ---- random-walk tables (imc_transport.jl:734-754, :786-797) ------------------------------------
  static double P_r(double a) {
    if (a == 0) return 1.0;
    double Pr = 0.0;
    for (int n = 1; n <= 100; ++n) {
      double pin = 3.141592653589793 * (double)n;
      double sgn = ((n - 1) & 1) ? -1.0 : 1.0;
      Pr += sgn * dm::exp_d(-a * (pin * pin)) * 2.0;
    }
    return Pr;
  }
  int rw_table(double lo, double hi, int n, double* a_out, double* pr_out, double* pt_out) override {
    IMC_RC(use_device());
    if (n < 2) { err = "rw_table: n < 2"; return IMC_ERR_ARG; }
    h_rw_a.resize(n); h_rw_pr.resize(n); h_rw_pt.resize(n);
    for (int i = 0; i < n; ++i) {
      double t = (double)i / (double)(n - 1);
      N a = N::from_d((1.0 - t) * lo + t * hi);          // T.(LinRange(lo, hi, n))
      N pr = N::from_d(P_r(a.d()));                       // stored into zeros(T)
      N pt = N::from_d(1.0) - pr;                         // T(1 - prVals[i])
      h_rw_a[i] = a.d(); h_rw_pr[i] = pr.d(); h_rw_pt[i] = pt.d();
      if (a_out) a_out[i] = a.d();
      if (pr_out) pr_out[i] = pr.d();
      if (pt_out) pt_out[i] = pt.d();
    }
    IMC_RC(upload(rw_a, h_rw_a.data(), n));
    IMC_RC(upload(rw_pt, h_rw_pt.data(), n));
    have_rw = true;
    return IMC_OK;
  }

  // ---- Update.update -------------------------------------------------------------------------
  int update(double dt) override {
    if (!have_mesh) { err = "update before set_mesh"; return IMC_ERR_STATE; }
    IMC_RC(use_device());
    k_update<P><<<grid_for(nc, 256), 256, 0, stream>>>(m, P::from_d(dt), cfg.linearized, cfg.marshak_quirk, temp_wide ? 1 : 0); ++n_launch;
    IMC_CK(cudaGetLastError());
    return IMC_OK;
  }

  // ---- Sourcing.sourcing ---------------------------------------------------------------------
  int source(double dt_, int64_t n_input, double cellmin_, int64_t step, int64_t n_census_global, imc_source_stats* out) override {
    if (!have_mesh) { err = "source before set_mesh"; return IMC_ERR_STATE; }
    if (step < 0 || step >= (1ll << 24)) { err = "time-step index outside [0, 2^24): particle ids are step << 40 | ordinal and the Philox counter carries the step in 28 bits"; return IMC_ERR_ARG; }
    IMC_RC(use_device());
    Cc dt = P::from_d(dt_), cellmin = P::from_d(cellmin_);
    alive_known = false;
    SrcArrays<P> s; s.e = src_e.p; s.q = src_q.p; s.ks = src_ks.p; s.cnt = src_cnt.p; s.nrg = src_nrg.p; s.q_em = src_qem.p;
    k_src_energies<P><<<grid_for(L.n_surf() + nc, 128), 128, 0, stream>>>(m, s, L, dt); ++n_launch;
    IMC_CK(cudaGetLastError());
    // totalenergy sums in the reference's association order
    if (geom == 1) {
      IMC_RC(jl_sum(s.q + L.body0(), nc, sums.p + 0));
      IMC_RC(jl_sum(s.q + L.rad0(), nc, sums.p + 1));
    } else {
      IMC_RC(jl_sum(s.q + 0, nx, sums.p + 0));
      IMC_RC(jl_sum(s.q + nx, nx, sums.p + 1));
      IMC_RC(jl_sum(s.q + 2 * nx, ny, sums.p + 2));
      IMC_RC(jl_sum(s.q + 2 * nx + ny, ny, sums.p + 3));
      IMC_RC(jl_sum(s.q + L.body0(), nc, sums.p + 4));
      IMC_RC(jl_sum(s.q + L.rad0(), nc, sums.p + 5));
    }
    IMC_RC(jl_sum(s.q_em, nc * ns, sums.p + 6));
    IMC_RC(jl_flush());
    long long n_census = n_census_global >= 0 ? n_census_global : n_part - n_holes;
    int wide_counts = (P::id == 0) && (std::max<int64_t>(n_input, cfg.n_max) > 65504);
    k_src_total<P><<<1, 1, 0, stream>>>(s, L, sums.p, src_sc.p, n_input, n_census, cfg.n_max, cellmin, wide_counts); ++n_launch;
    k_src_counts<P><<<grid_for(L.total(), 256), 256, 0, stream>>>(m, s, L, src_sc.p, cellmin, wide_counts); ++n_launch;
    IMC_CK(cudaGetLastError());
    IMC_RC(scan_counts(s.cnt, src_offs.p, L.total()));
    SrcScalars hsc; long long total = 0; Cc h_emsum = 0;
    gl_reset();
    gl_add(&src_sc.p->totalenergy, GK_RAW8); gl_add(&src_sc.p->nsrc, GK_RAW8); gl_add(&src_sc.p->bad, GK_I32);
    gl_add(scan_total.p, GK_RAW8); gl_add(sums.p + 6, GK_T);
    IMC_RC(gl_read());
    hsc.totalenergy = hpin[0]; hsc.nsrc = hpin[1]; hsc.bad = (int)hpin[2]; total = raw_ll(hpin[3]); h_emsum = (Cc)hpin[4];
    totalenergy = hsc.totalenergy;
    const long long world = cfg.world > 0 ? cfg.world : 1, rank = cfg.rank;
    long long n_local = total > rank ? (total - rank + world - 1) / world : 0;
    if (cfg.rng_mode == IMC_RNG_TAPE && n_local > 0 && total > st_slots) { err = "source tape has fewer slots than new particles"; return IMC_ERR_TAPE; }
    IMC_RC(ensure_capacity(n_part + n_local));
    if (n_local > 0) {
      IMC_CK(cudaMemsetAsync(over_flag.p, 0, sizeof(unsigned long long), stream));
      const long long n_blocks = (n_local + EMIT_THREADS - 1) / EMIT_THREADS;
      IMC_CK(src_block_entry.ensure((size_t)n_blocks + 1));
      k_src_block_entries<<<grid_for(n_blocks + 1, 256), 256, 0, stream>>>(src_offs.p, L.total(), n_local, (int)rank, (int)world, n_blocks, src_block_entry.p); ++n_launch;
      k_src_emit<P><<<(unsigned)n_blocks, EMIT_THREADS, 0, stream>>>(m, pb[cur].view(), s, L, src_offs.p, src_block_entry.p, n_part, n_local, (int)rank, (int)world, dt,
                                                                    rng_args(step, true), over_flag.p);
      ++n_launch;
      IMC_CK(cudaGetLastError());
      if (cfg.rng_mode == IMC_RNG_TAPE) {   // only a tape can run out
        unsigned long long over = 0;
        IMC_CK(cudaMemcpyAsync(&over, over_flag.p, sizeof over, cudaMemcpyDeviceToHost, stream));
        IMC_CK(cudaStreamSynchronize(stream));
        if (over) { err = "source tape exhausted"; return IMC_ERR_TAPE; }
      }
    }
    n_part += n_local;
    n_global_after_source = (n_census_global >= 0 ? (long long)n_census_global : n_part - n_local - n_holes) + total;
    if (out) {
      out->totalenergy = totalenergy; out->emitted_sum = (double)h_emsum; out->n_source = (int64_t)hsc.nsrc;
      out->n_new_global = total; out->n_new_local = n_local; out->n_particles = n_part - n_holes;
    }
    if (hsc.bad) { err = "non-finite particle count or unrepresentable energy (reference would throw)"; return IMC_ERR_NUMERIC; }
    return IMC_OK;
  }

  // ---- Transport -----------------------------------------------------------------------------
  int resolve_tally_mode() const {
    int mode = cfg.tally_mode;
    // AUTO: PAIRWISE = TRUE asks for the deterministic tree -> EXACT (Julia's pairwise order) while the deposit
    // records fit the budget, else the order-free fixed-point accumulation; PAIRWISE = FALSE -> float atomics
    // Float16 decks: sequential Float16 accumulation stagnates, i.e. the summation order is part of the
    // reference's result (it is what the mixed-precision study measures) -> EXACT as well, within the budget.
    // (Fixed-point accumulators would be faster wherever the tally fits in shared memory — integer shared-memory adds are
    // native on sm_100, Float32 / Float64 ones are compare-and-swap loops: Su-Olson tracking 2.57 vs 3.02 ms — but their
    // absolute quantum, 2^-62 of the largest possible sum, loses the cells 10+ decades below the maximum, which the
    // LINEARIZED temperature (a fourth root) exposes; AUTO therefore keeps float atomics for PAIRWISE = FALSE.)
    if (mode == IMC_TALLY_AUTO) mode = (cfg.pairwise || P::id == 0) ? IMC_TALLY_EXACT : IMC_TALLY_ATOMIC;
    return mode;
  }
  // smallest cell volume / scale, for fixed-point scaling
  double min_vol_h = 0, min_scale_h = 1, min_dx_h = 0;
  int compute_min_vol() {
    std::vector<double> hx(nx), hy(ny);
    IMC_RC(download(dx.p, nx, hx.data()));
    double mx = *std::min_element(hx.begin(), hx.end()), my = 1.0;
    if (geom == 2) { IMC_RC(download(dy.p, ny, hy.data())); my = *std::min_element(hy.begin(), hy.end()); }
    min_dx_h = mx; min_vol_h = mx * my;
    min_scale_h = 1e300;
    for (int k = 0; k < ns; ++k) min_scale_h = std::min(min_scale_h, (double)m.scales[k]);
    return IMC_OK;
  }
  static double pow2_floor_mul(double bound) {  // 2^S with bound * 2^S < 2^62
    if (!(bound > 0) || !std::isfinite(bound)) return 1.0;
    int e;
    std::frexp(bound, &e);  // bound < 2^e
    return std::ldexp(1.0, 62 - e);
  }
  // Fixed-point scales.  They must be identical on every rank (the host sums the integer buffers), so the
  // energy bound comes from replicated quantities: census energy at the end of the last step plus the
  // energy sourced this step, times the largest scale.  Without those (set_particles test flows) it falls
  // back to n * max(E), which is rank-local and therefore only valid for world == 1.
  double rad_total_h = 0;
  int prepare_fixed(TallyArgs& ta) {
    if (min_vol_h == 0) IMC_RC(compute_min_vol());
    double max_scale = 0;
    for (int k = 0; k < ns; ++k) max_scale = std::max(max_scale, (double)m.scales[k]);
    double tot = 2.0 * (rad_total_h + totalenergy) * max_scale;
    if (!(tot > 0)) {
      IMC_CK(cudaMemsetAsync(d_max.p, 0, sizeof(double), stream));
      k_max_energy<P><<<sm_count * 4, 256, 0, stream>>>(pb[cur].view(), n_part, d_max.p); ++n_launch;
      double maxE = 0;
      IMC_CK(cudaMemcpyAsync(&maxE, d_max.p, sizeof maxE, cudaMemcpyDeviceToHost, stream));
      IMC_CK(cudaStreamSynchronize(stream));
      tot = maxE * (double)std::max<long long>(n_part, 1) * (cfg.world > 0 ? cfg.world : 1);
    }
    fx_mul_dep = pow2_floor_mul(tot / min_vol_h);
    fx_mul_rad = pow2_floor_mul(tot / (min_vol_h * min_scale_h));
    // lostenergy keeps accumulating across transport calls until energycheck resets it: an integer already stored at the
    // previous scale is carried over to the new one (both are powers of two)
    const double new_lost = pow2_floor_mul(tot / min_scale_h);
    if (red_fixed && new_lost != fx_mul_lost) {
      k_rescale_fixed<<<1, 1, 0, stream>>>(reinterpret_cast<long long*>(red.p) + rb_sc0() + RB_LOST, new_lost / fx_mul_lost); ++n_launch;
    }
    fx_mul_lost = new_lost;
    ta.fx_mul = fx_mul_dep; ta.fx_mul_lost = fx_mul_lost;
    return IMC_OK;
  }
  size_t smem_for(int mode, long long nacc) const {
    size_t per = mode == IMC_TALLY_FIXED ? 8 : sizeof(typename AccType<P>::type);
    return (size_t)nacc * per;
  }
  // Dynamic shared memory of a launch stays within the 48 KB every kernel may request without an opt-in: `reserved` bytes
  // (the tracking kernels' counter slots, COUNTER_SMEM_BYTES) plus the accumulator sets.
  static constexpr size_t SMEM_LIMIT = 48 * 1024;
  static bool smem_fits(size_t one_set, size_t reserved) { return one_set + reserved <= SMEM_LIMIT; }
  // accumulator sets per block: as many as fit, at most one per warp of the block (a power of two)
  static int smem_copies(size_t one_set, size_t reserved) {
    int c = 1;
    while (c < TRACK_THREADS / 32 && one_set * (size_t)(2 * c) + reserved <= SMEM_LIMIT) c *= 2;
    return c;
  }


  // EXACT mode: records (key = tally cell, val) in reference order -> out[c] = reduction of cell c's records
  int exact_reduce_records(long long R, long long nacc, int pairwise, double* out) {
    IMC_CK(rec_start.ensure((size_t)nacc + 1));
    if (R > 0) {
      int end_bit = 1; while ((1ll << end_bit) < nacc + 1 && end_bit < 31) ++end_bit;
      size_t tb = 0;
      IMC_CK(exact_sort_pairs(nullptr, tb, rec_key[0].p, rec_key[1].p, rec_val[0].p, rec_val[1].p, R, end_bit, stream));
      IMC_CK(sort_temp.ensure(tb));
      IMC_CK(exact_sort_pairs(sort_temp.p, tb, rec_key[0].p, rec_key[1].p, rec_val[0].p, rec_val[1].p, R, end_bit, stream)); n_launch += 3;
    }
    k_exact_bounds<<<grid_for(nacc + 1, 256), 256, 0, stream>>>(rec_key[1].p, R, nacc, rec_start.p); ++n_launch;
    k_exact_reduce<P><<<grid_for(nacc, 128), 128, 0, stream>>>(rec_key[1].p, rec_val[1].p, rec_start.p, nacc, pairwise, out); ++n_launch;
    if (R >= EXACT_WARP_MIN) {   // cells with many records: a warp per cell, same order of additions
      unsigned grid = (unsigned)std::min<long long>((nacc * 32 + 255) / 256, (long long)sm_count * 8);
      // sequential sums skip their stagnant stretches (warp_seq_add_skip: same bits, 4x faster on the Float16 Su-Olson deck);
      // IMC_EXACT_SKIP=0 selects the plain chain for A/B runs
      static const int skip_stagnant = getenv("IMC_EXACT_SKIP") ? atoi(getenv("IMC_EXACT_SKIP")) : 1;
      k_exact_reduce_warp<P><<<grid, 256, 0, stream>>>(rec_key[1].p, rec_val[1].p, rec_start.p, nacc, pairwise, skip_stagnant, out); ++n_launch;
    }
    if (!pairwise && R >= EXACT_SEQBLOCK_MIN) {   // very long sequential segments: a block per cell, stagnant chunks skipped
      unsigned grid = (unsigned)std::min<long long>(nacc, (long long)sm_count * 2);
      k_exact_reduce_seqblock<P><<<grid, EXACT_BLOCK_THREADS, 0, stream>>>(rec_key[1].p, rec_val[1].p, rec_start.p, nacc, out); ++n_launch;
    }
    if (pairwise && R >= EXACT_BLOCK_MIN) {   // very long pairwise segments: a block per cell, leaves in parallel
      unsigned grid = (unsigned)std::min<long long>(nacc, (long long)sm_count * 2);
      k_exact_reduce_block<P><<<grid, EXACT_BLOCK_THREADS, 0, stream>>>(rec_val[1].p, rec_start.p, nacc, out); ++n_launch;
    }
    IMC_CK(cudaGetLastError());
    return IMC_OK;
  }
  template <int TK>
  void launch_event_tk(TrackArgs<P>& a, unsigned grid, size_t smem, long long n_active, int first) {
    if (geom == 1) k_track_event<P, 1, TK><<<grid, TRACK_THREADS, smem, stream>>>(a, n_active, first);
    else k_track_event<P, 2, TK><<<grid, TRACK_THREADS, smem, stream>>>(a, n_active, first);
  }
  void launch_event_kernel(TrackArgs<P>& a, unsigned grid, size_t smem, long long n_active, int first) {
    if (a.tally.mode == IMC_TALLY_ATOMIC && !a.tally.use_smem) launch_event_tk<TK_ATOMIC_G>(a, grid, smem, n_active, first);
    else if (a.tally.mode == IMC_TALLY_ATOMIC) launch_event_tk<TK_ATOMIC_S>(a, grid, smem, n_active, first);
    else if (a.tally.mode == IMC_TALLY_FIXED && !a.tally.use_smem) launch_event_tk<TK_FIXED_G>(a, grid, smem, n_active, first);
    else launch_event_tk<TK_FIXED_S>(a, grid, smem, n_active, first);
  }
  // event-based schedule: ev_batch segments per particle per launch, survivors compacted into the next launch's index list
  int launch_event(TrackArgs<P>& a, size_t smem) {
    static const int batch_env = getenv("IMC_EVENT_BATCH") ? atoi(getenv("IMC_EVENT_BATCH")) : 64;
    a.ev_batch = batch_env < 1 ? 1 : batch_env;
    IMC_CK(ev_list[0].ensure((size_t)n_part)); IMC_CK(ev_list[1].ensure((size_t)n_part));
    IMC_CK(ev_nseg.ensure((size_t)n_part)); IMC_CK(ev_extra.ensure((size_t)n_part)); IMC_CK(ev_count.ensure(1));
    a.ev_nseg = ev_nseg.p; a.ev_extra = ev_extra.p; a.ev_count = ev_count.p;
    long long n_active = n_part;
    int it = 0;
    smem += COUNTER_SMEM_BYTES;
    while (n_active > 0) {
      a.ev_in = it == 0 ? nullptr : ev_list[it & 1].p;
      a.ev_out = ev_list[(it + 1) & 1].p;
      IMC_CK(cudaMemsetAsync(ev_count.p, 0, sizeof(unsigned long long), stream));
      unsigned grid = (unsigned)std::min<long long>((n_active + TRACK_THREADS - 1) / TRACK_THREADS, (long long)sm_count * IMC_TRACK_MIN_BLOCKS);
      launch_event_kernel(a, grid, smem, n_active, it == 0);
      ++n_launch; ++it;
      IMC_CK(cudaGetLastError());
      unsigned long long cnt = 0;
      IMC_CK(cudaMemcpyAsync(&cnt, ev_count.p, sizeof cnt, cudaMemcpyDeviceToHost, stream));
      IMC_CK(cudaStreamSynchronize(stream));
      n_active = (long long)cnt;
    }
    ev_launches = it;
    return IMC_OK;
  }
  // history kernels: schedule x geometry x draw source x tally kind.  The Philox kernels with ATOMIC / FIXED tallies
  // are compiled per tally kind (no mode tests in the segment loop); EXACT passes and replay tapes read TallyArgs.
  template <bool TAPE, int TK>
  void launch_history(TrackArgs<P>& a, int variant, unsigned grid, size_t smem) {
    const bool rw = geom == 1 && cfg.randomwalk;
    if (variant == IMC_TRACK_REFILL) {
      if (rw) k_track_refill<P, 3, TAPE, TK><<<grid, TRACK_THREADS, smem, stream>>>(a);
      else if (geom == 1) k_track_refill<P, 1, TAPE, TK><<<grid, TRACK_THREADS, smem, stream>>>(a);
      else k_track_refill<P, 2, TAPE, TK><<<grid, TRACK_THREADS, smem, stream>>>(a);
    } else {
      if (rw) k_track1d_rw<P, TAPE, TK><<<grid, TRACK_THREADS, smem, stream>>>(a);
      else if (geom == 1) k_track1d<P, TAPE, TK><<<grid, TRACK_THREADS, smem, stream>>>(a);
      else k_track2d<P, TAPE, TK><<<grid, TRACK_THREADS, smem, stream>>>(a);
    }
  }
  int launch_track(TrackArgs<P>& a, int variant, unsigned grid, size_t smem) {
    if (variant == IMC_TRACK_EVENT) return launch_event(a, smem);
    IMC_CK(cudaMemsetAsync(over_flag.p + 1, 0, sizeof(unsigned long long), stream));
    smem += COUNTER_SMEM_BYTES;
    const bool tape = a.rng.tape != 0;
    if (tape) launch_history<true, TK_RUNTIME>(a, variant, grid, smem);
    else if (a.tally.mode == IMC_TALLY_ATOMIC && !a.tally.use_smem) launch_history<false, TK_ATOMIC_G>(a, variant, grid, smem);
    else if (a.tally.mode == IMC_TALLY_ATOMIC) launch_history<false, TK_ATOMIC_S>(a, variant, grid, smem);
    else if (a.tally.mode == IMC_TALLY_FIXED && !a.tally.use_smem) launch_history<false, TK_FIXED_G>(a, variant, grid, smem);
    else if (a.tally.mode == IMC_TALLY_FIXED) launch_history<false, TK_FIXED_S>(a, variant, grid, smem);
    else launch_history<false, TK_RUNTIME>(a, variant, grid, smem);
    ++n_launch;
    IMC_CK(cudaGetLastError());
    return IMC_OK;
  }

  int transport(double dt_, int64_t step, imc_transport_stats* out) override {
    if (!have_mesh) { err = "transport before set_mesh"; return IMC_ERR_STATE; }
    IMC_RC(use_device());
    if (step < 0 || step >= (1ll << 24)) { err = "time-step index outside [0, 2^24)"; return IMC_ERR_ARG; }
    if (cfg.randomwalk && !have_rw) { err = "random-walk tables not set (imc_rw_table)"; return IMC_ERR_STATE; }
    if (cfg.rng_mode == IMC_RNG_TAPE && n_part > tt_slots) { err = "transport tape has fewer slots than particles"; return IMC_ERR_TAPE; }
    if (n_part >= (1ll << 32)) { err = "more than 2^32 particles on one GPU"; return IMC_ERR_ARG; }
    alive_known = false; echeck_cached = false;
    int mode = resolve_tally_mode();
    if ((mode == IMC_TALLY_FIXED) != red_fixed) {  // representation change: start from a clean buffer
      IMC_CK(cudaMemsetAsync(red.p, 0, red_n * sizeof(double), stream));
      red_fixed = mode == IMC_TALLY_FIXED;
    }
    // mesh.energydep = zeros(...) (:45); per-call counters; lostenergy keeps accumulating
    IMC_CK(cudaMemsetAsync(red.p + rb_dep0(), 0, nc * ns * sizeof(double), stream));
    IMC_CK(cudaMemsetAsync(red.p + rb_sc0() + RB_SEG, 0, (RB_NSCALARS - RB_SEG) * sizeof(double), stream));
    IMC_CK(cudaMemsetAsync(over_flag.p, 0, 2 * sizeof(unsigned long long), stream));
    TrackArgs<P> a;
    a.m = m; a.p = pb[cur].view(); a.n = n_part; a.dt = P::from_d(dt_);
    a.rng = rng_args(step, false);
    a.tally.mode = mode; a.tally.nacc = (int)(nc * ns);
    a.tally.g_acc = red.p; a.tally.g_fx = reinterpret_cast<long long*>(red.p);
    a.tally.fx_mul = 1; a.tally.fx_mul_lost = 1; a.tally.sc0 = rb_sc0();
    a.tally.pass = 0; a.tally.rec_cnt = nullptr; a.tally.rec_off = nullptr; a.tally.rec_key = nullptr; a.tally.rec_val = nullptr; a.tally.lost_val = nullptr;
    if (mode == IMC_TALLY_FIXED) IMC_RC(prepare_fixed(a.tally));
    size_t smem = smem_for(mode, nc * ns);
    a.tally.use_smem = (smem_fits(smem, COUNTER_SMEM_BYTES) && mode != IMC_TALLY_EXACT) ? 1 : 0;
    a.tally.copies = a.tally.use_smem ? smem_copies(smem, COUNTER_SMEM_BYTES) : 1;
    smem = a.tally.use_smem ? smem * a.tally.copies : 0;
    static const int tforce = getenv("IMC_TALLY_ORDER") ? atoi(getenv("IMC_TALLY_ORDER")) : 0;   // 1: x fastest, 2: y fastest (experiments)
    dep_perm = geom == 2 && tforce == 2 && mode != IMC_TALLY_EXACT && !a.tally.use_smem;
    a.m.tsx = dep_perm ? ny : 1; a.m.tsy = dep_perm ? 1 : nx;
    // outcome records for replay checks (small populations only)
    bool record = n_part <= (1ll << 22) && n_holes == 0;   // list positions must be particle ordinals
    if (record) {
      IMC_CK(out_event.ensure((size_t)std::max<long long>(n_part, 1)));
      IMC_CK(out_nseg.ensure((size_t)std::max<long long>(n_part, 1)));
      a.out_event = out_event.p; a.out_nseg = out_nseg.p; out_n = n_part;
    } else { a.out_event = nullptr; a.out_nseg = nullptr; out_n = 0; }
    a.over_flag = over_flag.p;
    a.aVals = rw_a.p; a.ptVals = rw_pt.p; a.n_rw_table = (int)h_rw_a.size();
    // schedule: static grid-stride or dynamic warp refill.  AUTO measures both (alternating on the first
    // steps, re-probing every 32 calls) and keeps the one with the higher segments/s.
    int variant = cfg.track_mode;
    const bool event_ok = !(geom == 1 && cfg.randomwalk) && mode != IMC_TALLY_EXACT && cfg.rng_mode != IMC_RNG_TAPE && n_part < (1ll << 32);
    if (variant == IMC_TRACK_EVENT && !event_ok) variant = IMC_TRACK_AUTO;
    if (variant == IMC_TRACK_AUTO) {
      // measured selection: the two history schedules are probed on the first two calls and re-probed every
      // 32 / 256 calls; the event-based schedule (64 segments per particle per launch, survivors compacted) is probed once,
      // on the third call (IMC_AUTO_PROBE_EVENT=0 skips it).  Measured on B200: crooked pipe refill 73 ms, event 98 ms,
      // static 130 ms; Su-Olson static 2.6 ms, event 2.8 ms.  The fastest known variant runs in between.
      static const bool probe_event = !(getenv("IMC_AUTO_PROBE_EVENT") && atoi(getenv("IMC_AUTO_PROBE_EVENT")) == 0);
      // the losing schedule is re-measured every 32 calls while it is within 30 % of the winner, every 256 calls otherwise
      // (a lost probe costs one slow step: static is 1.7-2x slower than refill on the crooked pipe)
      const bool close = rate_static > 0 && rate_refill > 0 && std::min(rate_static, rate_refill) > 0.7 * std::max(rate_static, rate_refill);
      const long long period = (n_transport_calls < 2 || close) ? 32 : 256;
      long long phase = n_transport_calls % period;
      if (phase == 0) variant = IMC_TRACK_HISTORY;
      else if (phase == 1) variant = IMC_TRACK_REFILL;
      else if (n_transport_calls == 2 && event_ok && probe_event) variant = IMC_TRACK_EVENT;
      else {
        variant = rate_static >= rate_refill ? IMC_TRACK_HISTORY : IMC_TRACK_REFILL;
        if (event_ok && rate_event > rate_static && rate_event > rate_refill) variant = IMC_TRACK_EVENT;
      }
    }
    ++n_transport_calls;
    a.ev_in = nullptr; a.ev_out = nullptr; a.ev_count = nullptr; a.ev_nseg = nullptr; a.ev_extra = nullptr;
    a.queue = over_flag.p + 1;
    static const bool want_timeline = getenv("IMC_TRACK_TIMING") && atoi(getenv("IMC_TRACK_TIMING")) != 0;
    a.timeline = nullptr;
    if (want_timeline) {
      IMC_CK(timeline.ensure(4));
      const unsigned long long init[4] = {~0ull, ~0ull, 0ull, 0ull};
      IMC_CK(cudaMemcpyAsync(timeline.p, init, sizeof init, cudaMemcpyHostToDevice, stream));
      a.timeline = timeline.p;
    }
    a.refill_min = refill_min_env();
    a.queue_chunks = (unsigned long long)((n_part + QUEUE_CHUNK - 1) / QUEUE_CHUNK);   // tickets of the dynamic schedule's queue (k_track_refill)
    if (n_part > 0) {
      int blocks_per_sm = 2048 / TRACK_THREADS;
      if (smem > 0) blocks_per_sm = (int)std::max<size_t>(1, std::min<size_t>(blocks_per_sm, (200 * 1024) / (smem + COUNTER_SMEM_BYTES)));
      unsigned grid = (unsigned)std::min<long long>((n_part + TRACK_THREADS - 1) / TRACK_THREADS, (long long)sm_count * blocks_per_sm);
      IMC_CK(cudaEventRecord(ev0, stream));
      const long long budget = cfg.exact_record_budget > 0 ? cfg.exact_record_budget : (1ll << 28);
      const bool auto_mode = cfg.tally_mode == IMC_TALLY_AUTO;
      bool skip_exact = false;
      if (mode == IMC_TALLY_EXACT && auto_mode) {
        // AUTO: one deposit record per segment.  Predict this rank's records from replicated quantities — the segments all
        // ranks tracked in the previous step, at least one per particle of the global population — with a margin for the
        // growth from one step to the next, and go straight to the order-free accumulation when they cannot fit: the
        // count-only pass is a whole tracking pass (it used to run, and be thrown away, every step at scale), and every
        // rank must take the same branch because the host sums the buffers element by element.
        const double world_d = (double)(cfg.world > 0 ? cfg.world : 1);
        const double pred = std::max(last_global_segments, (double)std::max(n_global_after_source, n_part)) / world_d;
        skip_exact = 1.5 * pred > (double)budget;
      }
      if (mode == IMC_TALLY_EXACT && !skip_exact) {
        // pass 1: count the deposits of every particle (no side effects), scan -> record offsets
        IMC_CK(rec_cnt.ensure((size_t)n_part)); IMC_CK(rec_off.ensure((size_t)n_part + 1)); IMC_CK(lost_val.ensure((size_t)n_part));
        a.tally.pass = 1; a.tally.rec_cnt = rec_cnt.p;
        IMC_RC(launch_track(a, variant, grid, smem));
        IMC_RC(scan_counts(rec_cnt.p, rec_off.p, n_part));
        long long R = 0;
        IMC_CK(cudaMemcpyAsync(&R, scan_total.p, sizeof R, cudaMemcpyDeviceToHost, stream));
        IMC_CK(cudaStreamSynchronize(stream));
        if (R > budget) {
          if (cfg.tally_mode == IMC_TALLY_EXACT) { err = "EXACT tally mode: deposit records exceed exact_record_budget"; return IMC_ERR_NOMEM; }
          if (cfg.world > 1) {
            err = "AUTO tally mode: this rank's deposit records exceed exact_record_budget although the replicated prediction fitted; "
                  "every rank must accumulate in the same representation — set the tally mode explicitly or raise exact_record_budget";
            return IMC_ERR_NOMEM;
          }
          skip_exact = true;   // single GPU: fall back now (the prediction had no history yet)
        } else {
          for (int b = 0; b < 2; ++b) { IMC_CK(rec_key[b].ensure((size_t)std::max<long long>(R, 1))); IMC_CK(rec_val[b].ensure((size_t)std::max<long long>(R, 1))); }
          IMC_CK(cudaMemsetAsync(lost_val.p, 0xFF, (size_t)n_part * sizeof(double), stream));  // NaN = no loss
          a.tally.pass = 2; a.tally.rec_off = rec_off.p; a.tally.rec_key = rec_key[0].p; a.tally.rec_val = rec_val[0].p; a.tally.lost_val = lost_val.p;
          IMC_RC(launch_track(a, variant, grid, smem));
          IMC_RC(exact_reduce_records(R, nc * ns, cfg.pairwise, red.p + rb_dep0()));
          // vacuum losses: compact the per-particle slots in order, then add them as the reference's loop does
          k_lost_flags<<<grid_for(n_part, 256), 256, 0, stream>>>(lost_val.p, n_part, rec_cnt.p); ++n_launch;
          IMC_RC(scan_counts(rec_cnt.p, rec_off.p, n_part));
          long long n_lost = 0;
          IMC_CK(cudaMemcpyAsync(&n_lost, scan_total.p, sizeof n_lost, cudaMemcpyDeviceToHost, stream));
          IMC_CK(cudaStreamSynchronize(stream));
          if (n_lost > 0) {
            IMC_CK(lost_cval.ensure((size_t)n_lost)); IMC_CK(lost_cks.ensure((size_t)n_lost)); IMC_CK(lost_scratch.ensure((size_t)n_lost));
            k_lost_gather<<<grid_for(n_part, 256), 256, 0, stream>>>(lost_val.p, pb[cur].view().ks, rec_off.p, n_part, lost_cval.p, lost_cks.p); ++n_launch;
            k_exact_lost<P><<<1, 1, 0, stream>>>(lost_cval.p, lost_cks.p, n_lost, m, cfg.pairwise, lost_scratch.p, red.p + rb_sc0() + RB_LOST); ++n_launch;
          }
          IMC_CK(cudaGetLastError());
        }
      }
      if (mode == IMC_TALLY_EXACT && skip_exact) {   // AUTO: order-free fixed point (PAIRWISE) / float atomics instead
        mode = cfg.pairwise ? IMC_TALLY_FIXED : IMC_TALLY_ATOMIC;
        if (mode == IMC_TALLY_FIXED && !red_fixed) { IMC_CK(cudaMemsetAsync(red.p, 0, red_n * sizeof(double), stream)); red_fixed = true; }
        a.tally.mode = mode; a.tally.pass = 0;
        if (mode == IMC_TALLY_FIXED) IMC_RC(prepare_fixed(a.tally));
        smem = smem_for(mode, nc * ns); a.tally.use_smem = smem_fits(smem, COUNTER_SMEM_BYTES) ? 1 : 0;
        a.tally.copies = a.tally.use_smem ? smem_copies(smem, COUNTER_SMEM_BYTES) : 1; smem = a.tally.use_smem ? smem * a.tally.copies : 0;
        dep_perm = false; a.m.tsx = 1; a.m.tsy = nx;
        if (smem > 0) grid = (unsigned)std::min<long long>(grid, (long long)sm_count * std::max<size_t>(1, std::min<size_t>(2048 / TRACK_THREADS, (200 * 1024) / (smem + COUNTER_SMEM_BYTES))));
        IMC_RC(launch_track(a, variant, grid, smem));
      } else if (mode != IMC_TALLY_EXACT) {
        IMC_RC(launch_track(a, variant, grid, smem));
      }
      IMC_CK(cudaEventRecord(ev1, stream));
    }
    double sc[RB_NSCALARS];
    unsigned long long over = 0;
    gl_reset();
    for (int k = 0; k < RB_NSCALARS; ++k) gl_add(red.p + rb_sc0() + k, GK_RAW8);
    gl_add(over_flag.p, GK_RAW8);
    IMC_RC(gl_read());
    for (int k = 0; k < RB_NSCALARS; ++k) sc[k] = hpin[k];
    over = (unsigned long long)raw_ll(hpin[RB_NSCALARS]);
    float ms = 0;
    if (n_part > 0) IMC_CK(cudaEventElapsedTime(&ms, ev0, ev1));
    auto cnt = [&](int k) -> uint64_t {
      if (mode == IMC_TALLY_FIXED) { long long v; memcpy(&v, &sc[k], 8); return (uint64_t)v; }
      return (uint64_t)sc[k];
    };
    double lost;
    if (mode == IMC_TALLY_FIXED) { long long v; memcpy(&v, &sc[RB_LOST], 8); lost = (double)v / fx_mul_lost; } else lost = sc[RB_LOST];
    if (a.timeline) {
      unsigned long long tl[4];
      IMC_CK(cudaMemcpy(tl, timeline.p, sizeof tl, cudaMemcpyDeviceToHost));
      if (tl[0] != ~0ull && tl[2] != 0ull)
        fprintf(stderr, "[imc timeline] step %lld variant %d: kernel %.2f ms, queue empty after %.2f ms, tail %.2f ms, %lld particles\n",
                (long long)step, variant, (tl[2] - tl[0]) * 1e-6, tl[1] == ~0ull ? -1.0 : (tl[1] - tl[0]) * 1e-6,
                tl[1] == ~0ull ? -1.0 : (tl[2] - tl[1]) * 1e-6, n_part);
    }
    last_mode = mode;
    iterations += cnt(RB_SEG);
    if (ms > 0) {
      double rate = (double)cnt(RB_SEG) / ms;
      if (variant == IMC_TRACK_REFILL) rate_refill = rate; else if (variant == IMC_TRACK_EVENT) rate_event = rate; else rate_static = rate;
    }
    if (out) {
      out->lostenergy = (double)P::from_d(lost);
      out->segments = cnt(RB_SEG); out->segments_total = iterations; out->histories = (int64_t)cnt(RB_HIST);
      out->n_census = (int64_t)cnt(RB_CENSUS); out->n_absorbed = (int64_t)cnt(RB_ABSORBED); out->n_escaped = (int64_t)cnt(RB_ESCAPED);
      out->n_rw = (int64_t)cnt(RB_RW); out->n_errors = (int64_t)cnt(RB_ERRORS);
      out->variant = variant; out->tally_mode = mode; out->kernel_ms = ms;
    }
    if (over) { err = "transport tape exhausted"; return IMC_ERR_TAPE; }
    alive_known = true; alive_after_transport = (long long)cnt(RB_CENSUS);
    return IMC_OK;
  }

  // ---- Clean.clean ---------------------------------------------------------------------------
  // Lazy compaction.  Every kernel that walks the particle list skips the entries whose dead flag is set, so removing them is
  // only worth a pass over the whole population when enough of them have piled up: on the 10^8-particle Su-Olson decks a
  // handful of histories end per step, and the stable compaction that removes them copies all 10^8 survivors (1.9 of the step's
  // 5 ms).  clean() therefore always knows the survivors (that is length(particles) for the caller) and compacts when the dead
  // entries exceed 1/32 of the list; until then they stay as holes: n_part is the list length, n_part - n_holes the population.
  // Order, ids and therefore every result are unchanged.  Not used where list positions carry meaning: replay tapes (slot =
  // position), EXACT tallies (per-particle record counts) and lists small enough for outcome records (get_outcomes).
  long long n_holes = 0;
  static long long lazy_min_env() { const char* e = getenv("IMC_LAZY_CLEAN_MIN"); return e ? atoll(e) : (1ll << 22); }   // tests: 0 (always) / a huge value (never)
  bool lazy_clean_ok() const {
    return cfg.rng_mode != IMC_RNG_TAPE && resolve_tally_mode() != IMC_TALLY_EXACT && last_mode != IMC_TALLY_EXACT && n_part > lazy_min_env();
  }
  // The survivors of a transport call are its census outcomes (every tracked history ends in exactly one outcome, and the
  // other three set the dead flag), so clean() right after transport() knows length(particles) without counting: no
  // launch and no synchronisation when nothing has to move, and no synchronisation before the compaction otherwise.
  bool alive_known = false; long long alive_after_transport = 0;
  // what energycheck() needs, when the tally_finish() call before it has already read it (nothing may change radenergydens
  // or the lostenergy slot in between: every other entry point that could clears the flag)
  bool echeck_cached = false; Cc echeck_rad = 0; double echeck_lost_raw = 0;
  int count_blocks(long long blocks) {   // per-block survivor counts, scanned (device side only)
    IMC_CK(blk_cnt.ensure((size_t)blocks));
    IMC_CK(scan_total.ensure(1));
    k_alive_count<P><<<(unsigned)blocks, COMPACT_THREADS, 0, stream>>>(pb[cur].view(), n_part, geom, blk_cnt.p); ++n_launch;
    k_scan_small<<<1, 1024, 0, stream>>>(blk_cnt.p, blocks, scan_total.p); ++n_launch;
    IMC_CK(cudaGetLastError());
    return IMC_OK;
  }
  int count_alive(long long blocks, long long* total) {
    IMC_RC(count_blocks(blocks));
    gl_reset(); gl_add(scan_total.p, GK_RAW8);
    IMC_RC(gl_read());
    *total = raw_ll(hpin[0]);
    return IMC_OK;
  }
  int compact(long long blocks, long long total) {   // blk_cnt holds the scanned block counts of count_alive
    k_compact<P><<<(unsigned)blocks, COMPACT_THREADS, 0, stream>>>(pb[cur].view(), pb[cur ^ 1].view(), n_part, geom, blk_cnt.p); ++n_launch;
    IMC_CK(cudaGetLastError());
    n_part = total; n_holes = 0; alive_known = false;
    cur ^= 1;
    return IMC_OK;
  }
  int materialize() {   // before anything that exposes list positions
    if (n_holes == 0) return IMC_OK;
    const long long blocks = (n_part + COMPACT_TILE - 1) / COMPACT_TILE;
    long long total = 0;
    IMC_RC(count_alive(blocks, &total));
    return compact(blocks, total);
  }
  int clean(int64_t* n_alive) override {
    IMC_RC(use_device());
    if (n_part == 0) { n_holes = 0; if (n_alive) *n_alive = 0; return IMC_OK; }
    const long long blocks = (n_part + COMPACT_TILE - 1) / COMPACT_TILE;
    long long total = 0;
    const bool counted = !alive_known;
    if (alive_known) total = alive_after_transport; else IMC_RC(count_alive(blocks, &total));
    alive_known = false;
    if (total == n_part) { n_holes = 0; if (n_alive) *n_alive = n_part; return IMC_OK; }   // nobody died: the list is already compact
    if (lazy_clean_ok() && (n_part - total) * 32 <= n_part) { n_holes = n_part - total; if (n_alive) *n_alive = total; return IMC_OK; }
    if (!counted) IMC_RC(count_blocks(blocks));
    IMC_RC(compact(blocks, total));
    if (n_alive) *n_alive = n_part;
    return IMC_OK;
  }

  // ---- Tally.tally ---------------------------------------------------------------------------
  int tally_local() override {
    if (!have_mesh) { err = "tally before set_mesh"; return IMC_ERR_STATE; }
    IMC_RC(use_device());
    echeck_cached = false;
    int mode = red_fixed ? IMC_TALLY_FIXED : (last_mode == IMC_TALLY_EXACT || (resolve_tally_mode() == IMC_TALLY_EXACT && n_transport_calls == 0) ? IMC_TALLY_EXACT : IMC_TALLY_ATOMIC);
    if (mode == IMC_TALLY_EXACT && n_part > (cfg.exact_record_budget > 0 ? cfg.exact_record_budget : (1ll << 28))) mode = IMC_TALLY_ATOMIC;
    IMC_CK(cudaMemsetAsync(red.p + rb_rad0(), 0, nc * sizeof(double), stream));
    if (n_part == 0) return IMC_OK;
    TallyArgs ta;
    ta.pass = 0; ta.rec_cnt = nullptr; ta.rec_off = nullptr; ta.rec_key = nullptr; ta.rec_val = nullptr; ta.lost_val = nullptr;
    if (mode == IMC_TALLY_EXACT) {  // per-cell vectors + Julia sum, in particle order (imc_tally.jl:84-113, Q19)
      for (int b = 0; b < 2; ++b) { IMC_CK(rec_key[b].ensure((size_t)n_part)); IMC_CK(rec_val[b].ensure((size_t)n_part)); }
      ta.mode = mode; ta.nacc = (int)nc; ta.use_smem = 0; ta.copies = 1; ta.g_acc = nullptr; ta.g_fx = nullptr; ta.fx_mul = 1; ta.fx_mul_lost = 1; ta.sc0 = 0;
      ta.pass = 2; ta.rec_key = rec_key[0].p; ta.rec_val = rec_val[0].p;
      k_census_tally<P><<<grid_for(n_part, TRACK_THREADS), TRACK_THREADS, 0, stream>>>(m, pb[cur].view(), n_part, ta); ++n_launch;
      IMC_CK(cudaGetLastError());
      return exact_reduce_records(n_part, nc, 1, red.p + rb_rad0());
    }
    ta.mode = mode; ta.nacc = (int)nc; ta.g_acc = red.p + rb_rad0(); ta.g_fx = reinterpret_cast<long long*>(red.p) + rb_rad0();
    if (mode == IMC_TALLY_FIXED && fx_mul_rad == 1) IMC_RC(prepare_fixed(ta));
    ta.fx_mul = fx_mul_rad; ta.fx_mul_lost = fx_mul_lost; ta.sc0 = 0;
    size_t smem = smem_for(mode, nc);
    ta.use_smem = smem_fits(smem, 0) ? 1 : 0;   // (global RED instead, measured on the 10^8-particle Su-Olson deck: 27.7 against 3.8 ms per step)
    ta.copies = ta.use_smem ? smem_copies(smem, 0) : 1;
    smem = ta.use_smem ? smem * ta.copies : 0;
    int blocks_per_sm = 2048 / TRACK_THREADS;
    const long long groups = (n_part + 3) / 4;   // a thread takes four consecutive particles
    unsigned grid = (unsigned)std::min<long long>((groups + TRACK_THREADS - 1) / TRACK_THREADS, (long long)sm_count * blocks_per_sm);
    k_census_tally<P><<<grid, TRACK_THREADS, smem, stream>>>(m, pb[cur].view(), n_part, ta); ++n_launch;
    IMC_CK(cudaGetLastError());
    return IMC_OK;
  }
  int tally_finish(double t_, double dt_, imc_tally_stats* out) override {
    if (!have_mesh) { err = "tally before set_mesh"; return IMC_ERR_STATE; }
    IMC_RC(use_device());
    k_acc_to_field<P><<<grid_for(nc * ns, 256), 256, 0, stream>>>(red.p + rb_dep0(), red_fixed ? 1 : 0, fx_mul_dep, nc * ns, energydep.p,
                                                                     nc, dep_perm ? nx : 0, ny, 1); ++n_launch;
    k_acc_to_field<P><<<grid_for(nc, 256), 256, 0, stream>>>(red.p + rb_rad0(), red_fixed ? 1 : 0, fx_mul_rad, nc, radenergydens.p, nc, 0, 0, 0); ++n_launch;
    TallyScratch<P> s; s.q_dep = q_dep.p; s.q_tot = q_tot.p; s.q_rad = q_rad.p;
    k_tally_finish<P><<<grid_for(nc, 256), 256, 0, stream>>>(m, s, P::from_d(dt_), t_ == 0.0 ? 1 : 0, cfg.linearized, temp_wide ? 1 : 0); ++n_launch;
    IMC_CK(cudaGetLastError());
    if (cfg.linearized && P::id != 2) temp_wide = true;
    // per-plane Julia sums of (energydep .* vol) ./ scale (device slots 16 ..), sum(nrg_inc), sum(matenergydens + radenergydens),
    // the radiation energy: one batch, one readback below
    for (int k = 0; k < ns; ++k) IMC_RC(jl_sum(q_dep.p + nc * k, nc, sums.p + 16 + k));
    IMC_RC(jl_sum(nrg_inc.p, nc, sums.p + 12));
    IMC_RC(jl_sum(q_tot.p, nc, sums.p + 13));
    k_rad_energy<P><<<grid_for(nc, 256), 256, 0, stream>>>(m, q_rad.p); ++n_launch;
    IMC_RC(jl_sum(q_rad.p, nc, sums.p + 14));
    IMC_RC(jl_flush());
    double ninf = -INFINITY;
    IMC_CK(cudaMemcpyAsync(d_max.p, &ninf, sizeof ninf, cudaMemcpyHostToDevice, stream));
    IMC_CK(cudaMemsetAsync(d_flag.p, 0, sizeof(int), stream));
    k_max_f64<<<sm_count * 2, 256, 0, stream>>>(temp.p, nc, d_max.p, d_flag.p); ++n_launch;
    IMC_CK(cudaGetLastError());
    Cc h2[3]; double mx; int has_nan; double gseg = 0;
    std::vector<Cc> plane(ns);
    gl_reset();
    for (int k = 0; k < 3; ++k) gl_add(sums.p + 12 + k, GK_T);
    gl_add(d_max.p, GK_RAW8); gl_add(d_flag.p, GK_I32);
    gl_add(red.p + rb_sc0() + RB_SEG, GK_RAW8);                                     // summed over ranks by the host
    for (int k = 0; k < ns; ++k) gl_add(sums.p + 16 + k, GK_T);
    const int i_lost = gl_add(red.p + rb_sc0() + RB_LOST, GK_RAW8);                 // for energycheck (below)
    IMC_RC(gl_read());
    for (int k = 0; k < 3; ++k) h2[k] = (Cc)hpin[k];
    mx = hpin[3]; has_nan = (int)hpin[4]; gseg = hpin[5];
    for (int k = 0; k < ns; ++k) plane[k] = (Cc)hpin[6 + k];
    // EnergyCheck.energychecker needs sum(radenergydens .* volume) and lostenergy: the first was just formed for this
    // stage's own statistics and the second sits in the buffer read above, so the call that follows Tally.tally in the
    // reference's loop needs no launch and no synchronisation of its own
    if (i_lost >= 0) { echeck_rad = h2[2]; echeck_lost_raw = hpin[i_lost]; echeck_cached = true; }
    N ted;
    for (int k = 0; k < ns; ++k) ted = ted + N(plane[k]);                            // :51 / :55
    totalenergydep = ted.d();
    rad_total_h = (double)h2[2];
    if (red_fixed) { long long v; memcpy(&v, &gseg, 8); last_global_segments = (double)v; } else last_global_segments = gseg;
    IMC_RC(history_push());
    if (out) {
      out->totalenergydep = totalenergydep; out->energy_increase = (double)h2[0];
      out->max_temp = has_nan ? NAN : mx; out->total_energy_density = (double)h2[1];
    }
    return IMC_OK;
  }

  // ---- EnergyCheck.energychecker ---------------------------------------------------------------
  int energycheck(imc_energy_stats* out) override {
    if (!have_mesh) { err = "energycheck before set_mesh"; return IMC_ERR_STATE; }
    IMC_RC(use_device());
    Cc h; double lost_raw;
    if (echeck_cached) { h = echeck_rad; lost_raw = echeck_lost_raw; }   // read by the tally_finish call just before
    else {
      k_rad_energy<P><<<grid_for(nc, 256), 256, 0, stream>>>(m, q_rad.p); ++n_launch;
      IMC_CK(cudaGetLastError());
      IMC_RC(jl_sum(q_rad.p, nc, sums.p + 14));
      IMC_RC(jl_flush());
      gl_reset(); gl_add(sums.p + 14, GK_T); gl_add(red.p + rb_sc0() + RB_LOST, GK_RAW8);
      IMC_RC(gl_read());
      h = (Cc)hpin[0]; lost_raw = hpin[1];
    }
    echeck_cached = false;
    double lost;
    if (red_fixed) { long long v; memcpy(&v, &lost_raw, 8); lost = (double)v / fx_mul_lost; } else lost = lost_raw;
    N radenergy(h), te = N::from_d(totalenergy), ted = N::from_d(totalenergydep), old = N::from_d(radenergyold), lo = N::from_d(lost);
    N change = radenergy - old;
    N e = (((te - ted) - change) - lo) / te;                                          // :34
    if (out) { out->radenergy = radenergy.d(); out->radenergy_change = change.d(); out->lostenergy = lo.d(); out->energy_error = e.d(); }
    radenergyold = radenergy.d();                                                     // :36
    IMC_CK(cudaMemsetAsync(red.p + rb_sc0() + RB_LOST, 0, sizeof(double), stream));   // :37
    return IMC_OK;
  }

  int reduce_buffer(void** ptr, int64_t* n, int32_t* is_int) override {
    if (!have_mesh) { err = "reduce_buffer before set_mesh"; return IMC_ERR_STATE; }
    IMC_RC(use_device());
    IMC_CK(cudaStreamSynchronize(stream));  // the host's collective runs on another stream
    *ptr = red.p; *n = red_n; *is_int = red_fixed ? 1 : 0;
    return IMC_OK;
  }

  int get_field(int f, double* dst, int64_t n) override {
    if (!have_mesh) { err = "get_field before set_mesh"; return IMC_ERR_STATE; }
    IMC_RC(use_device());
    const S* src = nullptr; long long len = nc;
    switch (f) {
      case IMC_FIELD_TEMP:
        if (n != nc) { err = "get_field: size"; return IMC_ERR_ARG; }
        IMC_CK(cudaMemcpyAsync(dst, temp.p, nc * sizeof(double), cudaMemcpyDeviceToHost, stream));
        IMC_CK(cudaStreamSynchronize(stream));
        return IMC_OK;
      case IMC_FIELD_FLECK: src = fleck.p; break;
      case IMC_FIELD_BETA: src = beta.p; break;
      case IMC_FIELD_BEE: src = bee.p; break;
      case IMC_FIELD_SIGMA_A: src = sa.p; break;
      case IMC_FIELD_SIGMA_S: src = ss.p; break;
      case IMC_FIELD_ENERGYDEP: src = energydep.p; len = nc * ns; break;
      case IMC_FIELD_EMITTEDENERGY: src = emittedenergy.p; len = nc * ns; break;
      case IMC_FIELD_MATENERGYDENS: src = matenergydens.p; break;
      case IMC_FIELD_RADENERGYDENS: src = radenergydens.p; break;
      case IMC_FIELD_NRG_INC: src = nrg_inc.p; break;
      default: err = "get_field: unknown field"; return IMC_ERR_ARG;
    }
    if (n != len) { err = "get_field: size"; return IMC_ERR_ARG; }
    return download(src, (size_t)len, dst);
  }
  // ---- native-precision transfers (no Float64 staging): the field's device buffer <-> the host's Array{T} ----
  const S* field_ptr(int f, long long* len) {
    *len = nc;
    switch (f) {
      case IMC_FIELD_FLECK: return fleck.p;
      case IMC_FIELD_BETA: return beta.p;
      case IMC_FIELD_BEE: return bee.p;
      case IMC_FIELD_SIGMA_A: return sa.p;
      case IMC_FIELD_SIGMA_S: return ss.p;
      case IMC_FIELD_ENERGYDEP: *len = nc * ns; return energydep.p;
      case IMC_FIELD_EMITTEDENERGY: *len = nc * ns; return emittedenergy.p;
      case IMC_FIELD_MATENERGYDENS: return matenergydens.p;
      case IMC_FIELD_RADENERGYDENS: return radenergydens.p;
      case IMC_FIELD_NRG_INC: return nrg_inc.p;
      default: return nullptr;
    }
  }
  int field_elsize(int f) override {
    if (f < 0 || f >= IMC_FIELD_COUNT_) return 0;
    return (f == IMC_FIELD_TEMP && temp_wide) ? 8 : (int)sizeof(S);
  }
  DBuf<S> stage_t;  // device staging for mesh.temp (kept as its Float64 image on the device) in T
  int get_field_native(int f, void* dst, int64_t bytes) override {
    if (!have_mesh) { err = "get_field before set_mesh"; return IMC_ERR_STATE; }
    IMC_RC(use_device());
    if (f == IMC_FIELD_TEMP) {
      if (bytes != nc * field_elsize(f)) { err = "get_field_native: size"; return IMC_ERR_ARG; }
      if (temp_wide) IMC_CK(cudaMemcpyAsync(dst, temp.p, (size_t)bytes, cudaMemcpyDefault, stream));
      else {
        IMC_CK(stage_t.ensure((size_t)nc));
        k_from_f64<P><<<grid_for(nc, 256), 256, 0, stream>>>(temp.p, nc, stage_t.p); ++n_launch;
        IMC_CK(cudaMemcpyAsync(dst, stage_t.p, (size_t)bytes, cudaMemcpyDefault, stream));
      }
      IMC_CK(cudaStreamSynchronize(stream));
      return IMC_OK;
    }
    long long len = 0;
    const S* src = field_ptr(f, &len);
    if (!src) { err = "get_field_native: unknown field"; return IMC_ERR_ARG; }
    if (bytes != len * (long long)sizeof(S)) { err = "get_field_native: size"; return IMC_ERR_ARG; }
    IMC_CK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, stream));
    IMC_CK(cudaStreamSynchronize(stream));
    return IMC_OK;
  }
  int set_state_native(const void* temp_, const void* mat, const void* rad) override {
    if (!have_mesh) { err = "set_state before set_mesh"; return IMC_ERR_STATE; }
    IMC_RC(use_device());
    echeck_cached = false;
    if (temp_) {
      if (temp_wide) IMC_CK(cudaMemcpyAsync(temp.p, temp_, nc * sizeof(double), cudaMemcpyDefault, stream));
      else {
        IMC_CK(stage_t.ensure((size_t)nc));
        IMC_CK(cudaMemcpyAsync(stage_t.p, temp_, nc * sizeof(S), cudaMemcpyDefault, stream));
        k_to_f64<P><<<grid_for(nc, 256), 256, 0, stream>>>(stage_t.p, nc, temp.p); ++n_launch;
      }
    }
    if (mat) IMC_CK(cudaMemcpyAsync(matenergydens.p, mat, nc * sizeof(S), cudaMemcpyDefault, stream));
    if (rad) IMC_CK(cudaMemcpyAsync(radenergydens.p, rad, nc * sizeof(S), cudaMemcpyDefault, stream));
    IMC_CK(cudaGetLastError());
    IMC_CK(cudaStreamSynchronize(stream));
    return IMC_OK;
  }
  void* stream_handle() override { return (void*)stream; }

  // ---- per-step history on the device (imc_tally.jl:58, :138-142) ---------------------------------------------
  DBuf<double> hist_temp; DBuf<S> hist_mat, hist_rad, hist_inc;
  long long hist_cap = 0, hist_n = 0, hist_dropped = 0;
  int history_enable(int64_t cap) override {
    if (!have_mesh) { err = "history_enable before set_mesh"; return IMC_ERR_STATE; }
    if (cap < 0) { err = "history_enable: negative capacity"; return IMC_ERR_ARG; }
    IMC_RC(use_device());
    IMC_CK(cudaStreamSynchronize(stream));
    hist_temp.release(); hist_mat.release(); hist_rad.release(); hist_inc.release();
    hist_cap = hist_n = hist_dropped = 0;
    if (cap > 0) {
      const size_t n = (size_t)cap * (size_t)nc;
      IMC_CK(hist_temp.alloc(n, false)); IMC_CK(hist_mat.alloc(n, false)); IMC_CK(hist_rad.alloc(n, false)); IMC_CK(hist_inc.alloc(n, false));
      hist_cap = cap;
    }
    return IMC_OK;
  }
  int history_push() {   // end of Tally.tally
    if (hist_cap == 0) return IMC_OK;
    if (hist_n == hist_cap) { ++hist_dropped; return IMC_OK; }
    const size_t off = (size_t)hist_n * (size_t)nc;
    IMC_CK(cudaMemcpyAsync(hist_temp.p + off, temp.p, nc * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    IMC_CK(cudaMemcpyAsync(hist_mat.p + off, matenergydens.p, nc * sizeof(S), cudaMemcpyDeviceToDevice, stream));
    IMC_CK(cudaMemcpyAsync(hist_rad.p + off, radenergydens.p, nc * sizeof(S), cudaMemcpyDeviceToDevice, stream));
    IMC_CK(cudaMemcpyAsync(hist_inc.p + off, nrg_inc.p, nc * sizeof(S), cudaMemcpyDeviceToDevice, stream));
    ++hist_n;
    return IMC_OK;
  }
  int history_count(int64_t* stored, int64_t* dropped) override {
    if (stored) *stored = hist_n;
    if (dropped) *dropped = hist_dropped;
    return IMC_OK;
  }
  int history_get(int f, int64_t first, int64_t count, void* dst, int64_t bytes) override {
    IMC_RC(use_device());
    if (first < 0 || count < 0 || first + count > hist_n) { err = "history_get: snapshot range outside the stored history"; return IMC_ERR_ARG; }
    const void* src = nullptr; size_t es = sizeof(S);
    switch (f) {
      case IMC_FIELD_TEMP: src = hist_temp.p; es = sizeof(double); break;
      case IMC_FIELD_MATENERGYDENS: src = hist_mat.p; break;
      case IMC_FIELD_RADENERGYDENS: src = hist_rad.p; break;
      case IMC_FIELD_NRG_INC: src = hist_inc.p; break;
      default: err = "history_get: field has no history (temp, matenergydens, radenergydens, nrg_inc)"; return IMC_ERR_ARG;
    }
    if (bytes != (int64_t)(count * nc * (long long)es)) { err = "history_get: size"; return IMC_ERR_ARG; }
    if (count == 0) return IMC_OK;
    IMC_CK(cudaMemcpyAsync(dst, static_cast<const char*>(src) + (size_t)first * (size_t)nc * es, (size_t)bytes, cudaMemcpyDeviceToHost, stream));
    IMC_CK(cudaStreamSynchronize(stream));
    return IMC_OK;
  }
  int history_clear() override { hist_n = 0; hist_dropped = 0; return IMC_OK; }

  int set_state(const double* temp_, const double* mat, const double* rad) override {
    if (!have_mesh) { err = "set_state before set_mesh"; return IMC_ERR_STATE; }
    IMC_RC(use_device());
    echeck_cached = false;
    if (temp_) {
      if (temp_wide) IMC_CK(cudaMemcpyAsync(temp.p, temp_, nc * sizeof(double), cudaMemcpyHostToDevice, stream));
      else {  // round through T, keep the Float64 image
        IMC_CK(stage.ensure((size_t)nc));
        IMC_CK(cudaMemcpyAsync(stage.p, temp_, nc * sizeof(double), cudaMemcpyHostToDevice, stream));
        k_round_f64<P><<<grid_for(nc, 256), 256, 0, stream>>>(stage.p, nc, temp.p); ++n_launch;
      }
    }
    if (mat) IMC_RC(upload_into(matenergydens.p, mat, (size_t)nc));
    if (rad) IMC_RC(upload_into(radenergydens.p, rad, (size_t)nc));
    IMC_CK(cudaStreamSynchronize(stream));
    return IMC_OK;
  }

  int64_t num_particles() override { return n_part - n_holes; }
  int64_t launches() override { return n_launch; }
  int get_particles(double* slots, uint64_t* ids, int64_t capacity) override {
    IMC_RC(use_device());
    IMC_RC(materialize());
    if (capacity < n_part) { err = "get_particles: capacity"; return IMC_ERR_ARG; }
    if (n_part == 0) return IMC_OK;
    int nsl = geom == 1 ? 9 : 10;
    DBuf<double> d_slots; DBuf<unsigned long long> d_ids;
    IMC_CK(d_slots.alloc((size_t)n_part * nsl, false));
    IMC_CK(d_ids.alloc((size_t)n_part, false));
    k_export_particles<P><<<grid_for(n_part, 256), 256, 0, stream>>>(m, pb[cur].view(), n_part, d_slots.p, d_ids.p); ++n_launch;
    IMC_CK(cudaGetLastError());
    IMC_CK(cudaMemcpyAsync(slots, d_slots.p, (size_t)n_part * nsl * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (ids) IMC_CK(cudaMemcpyAsync(ids, d_ids.p, (size_t)n_part * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    IMC_CK(cudaStreamSynchronize(stream));
    return IMC_OK;
  }
  int set_particles(const double* slots, const uint64_t* ids, int64_t n) override {
    if (!have_mesh) { err = "set_particles before set_mesh"; return IMC_ERR_STATE; }
    IMC_RC(use_device());
    if (n < 0) { err = "set_particles: n < 0"; return IMC_ERR_ARG; }
    n_part = 0; n_holes = 0; alive_known = false;
    IMC_RC(ensure_capacity(n));
    if (n == 0) return IMC_OK;
    int nsl = geom == 1 ? 9 : 10;
    DBuf<double> d_slots; DBuf<unsigned long long> d_ids;
    IMC_CK(d_slots.alloc((size_t)n * nsl, false));
    IMC_CK(cudaMemcpyAsync(d_slots.p, slots, (size_t)n * nsl * sizeof(double), cudaMemcpyHostToDevice, stream));
    if (ids) { IMC_CK(d_ids.alloc((size_t)n, false)); IMC_CK(cudaMemcpyAsync(d_ids.p, ids, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, stream)); }
    IMC_CK(cudaMemsetAsync(d_flag.p, 0, sizeof(int), stream));
    k_import_particles<P><<<grid_for(n, 256), 256, 0, stream>>>(m, pb[cur].view(), n, d_slots.p, ids ? d_ids.p : nullptr, d_flag.p); ++n_launch;
    IMC_CK(cudaGetLastError());
    int bad = 0;
    IMC_CK(cudaMemcpyAsync(&bad, d_flag.p, sizeof bad, cudaMemcpyDeviceToHost, stream));
    IMC_CK(cudaStreamSynchronize(stream));
    if (bad) { err = bad & 2 ? "set_particles: cell index out of range" : "set_particles: energyscale not in ENERGYSCALES"; return IMC_ERR_ARG; }
    n_part = n;
    return IMC_OK;
  }
  int set_transport_tape(const double* u, int nu, const double* e, int ne, int64_t slots) override {
    IMC_RC(use_device());
    if (nu < 0 || ne < 0 || slots < 0) { err = "tape: negative size"; return IMC_ERR_ARG; }
    IMC_CK(tt_uni.alloc((size_t)nu * slots, false)); IMC_CK(tt_exp.alloc((size_t)ne * slots, false));
    if ((size_t)nu * slots) IMC_CK(cudaMemcpy(tt_uni.p, u, (size_t)nu * slots * sizeof(double), cudaMemcpyHostToDevice));
    if ((size_t)ne * slots) IMC_CK(cudaMemcpy(tt_exp.p, e, (size_t)ne * slots * sizeof(double), cudaMemcpyHostToDevice));
    tt_nuni = nu; tt_nexp = ne; tt_slots = slots;
    return IMC_OK;
  }
  int set_source_tape(const double* u, int nu, int64_t slots) override {
    IMC_RC(use_device());
    if (nu < 0 || slots < 0) { err = "tape: negative size"; return IMC_ERR_ARG; }
    IMC_CK(st_uni.alloc((size_t)nu * slots, false));
    if ((size_t)nu * slots) IMC_CK(cudaMemcpy(st_uni.p, u, (size_t)nu * slots * sizeof(double), cudaMemcpyHostToDevice));
    st_nuni = nu; st_slots = slots;
    return IMC_OK;
  }
  // ---- Sourcing.sample_planck (imc_sourcing.jl:372-399) --------------------------------------------------
  int sample_planck(int64_t n, int64_t step, double* out) override {
    IMC_RC(use_device());
    if (n < 0 || (n > 0 && !out)) { err = "sample_planck: bad arguments"; return IMC_ERR_ARG; }
    if (n == 0) return IMC_OK;
    if (cfg.rng_mode == IMC_RNG_TAPE && n > st_slots) { err = "source tape has fewer slots than samples"; return IMC_ERR_TAPE; }
    IMC_CK(stage.ensure((size_t)n));
    IMC_CK(over_flag.ensure(2));
    IMC_CK(cudaMemsetAsync(over_flag.p, 0, sizeof(unsigned long long), stream));
    k_sample_planck<P><<<grid_for(n, 128), 128, 0, stream>>>(rng_args(step, true), n, stage.p, over_flag.p); ++n_launch;
    IMC_CK(cudaGetLastError());
    unsigned long long over = 0;
    IMC_CK(cudaMemcpyAsync(out, stage.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    IMC_CK(cudaMemcpyAsync(&over, over_flag.p, sizeof over, cudaMemcpyDeviceToHost, stream));
    IMC_CK(cudaStreamSynchronize(stream));
    if (over) { err = "source tape exhausted"; return IMC_ERR_TAPE; }
    return IMC_OK;
  }
  // ---- restart point in device memory (include/imc.h imc_checkpoint) ------------------------------------------
  struct CkptHost {
    long long n_part, n_holes; bool temp_wide, red_fixed, dep_perm; double fx_mul_dep, fx_mul_rad, fx_mul_lost, rad_total_h;
    int last_mode; double totalenergy, totalenergydep, radenergyold; uint64_t iterations; long long n_transport_calls;
    double rate_static, rate_refill, rate_event; long long hist_n, hist_dropped; double last_global_segments; long long n_global_after_source;
  };
  DBuf<unsigned char> ckpt_blob;
  CkptHost ckpt_host;
  bool ckpt_valid = false;
  // every device array a step can change, as (address, bytes); particles: the live prefix of the current buffer
  std::vector<std::pair<void*, size_t>> ckpt_items(long long np) {
    std::vector<std::pair<void*, size_t>> v;
    auto add = [&](void* p, size_t b) { if (p && b) v.emplace_back(p, b); };
    const size_t S_ = sizeof(S);
    add(sa.p, nc * S_); add(ss.p, nc * S_); add(fleck.p, nc * S_); add(beta.p, nc * S_); add(bee.p, nc * S_);
    add(temp.p, nc * sizeof(double)); add(matenergydens.p, nc * S_); add(radenergydens.p, nc * S_); add(nrg_inc.p, nc * S_);
    add(energydep.p, nc * ns * S_); add(emittedenergy.p, nc * ns * S_);
    if (geom == 1) add(cp1.p, nc * sizeof(CellProp1<P>)); else add(cp2.p, nc * sizeof(CellProp2<P>));
    add(red.p, (size_t)red_n * sizeof(double));
    PartBufs<P>& b = pb[cur];
    add(b.t.p, np * S_); add(b.x.p, np * S_); add(b.mu.p, np * S_); add(b.E.p, np * S_); add(b.E0.p, np * S_);
    add(b.cx.p, np * sizeof(int)); add(b.ks.p, (size_t)np); add(b.id.p, np * sizeof(unsigned long long));
    if (geom == 2) { add(b.y.p, np * S_); add(b.cy.p, np * sizeof(int)); } else add(b.origin.p, np * sizeof(int));
    return v;
  }
  int checkpoint(int op) override {
    if (!have_mesh) { err = "checkpoint before set_mesh"; return IMC_ERR_STATE; }
    IMC_RC(use_device());
    if (op == IMC_CKPT_DROP) { IMC_CK(cudaStreamSynchronize(stream)); ckpt_blob.release(); ckpt_valid = false; return IMC_OK; }
    if (op == IMC_CKPT_SAVE) {
      auto items = ckpt_items(n_part);
      size_t total = 0;
      for (auto& it : items) total += (it.second + 255) & ~(size_t)255;
      IMC_CK(ckpt_blob.ensure(total));
      size_t off = 0;
      for (auto& it : items) { IMC_CK(cudaMemcpyAsync(ckpt_blob.p + off, it.first, it.second, cudaMemcpyDeviceToDevice, stream)); off += (it.second + 255) & ~(size_t)255; }
      ckpt_host = CkptHost{n_part, n_holes, temp_wide, red_fixed, dep_perm, fx_mul_dep, fx_mul_rad, fx_mul_lost, rad_total_h, last_mode, totalenergy,
                           totalenergydep, radenergyold, iterations, n_transport_calls, rate_static, rate_refill, rate_event, hist_n, hist_dropped, last_global_segments, n_global_after_source};
      IMC_CK(cudaStreamSynchronize(stream));
      ckpt_valid = true;
      return IMC_OK;
    }
    if (op != IMC_CKPT_RESTORE) { err = "checkpoint: unknown op"; return IMC_ERR_ARG; }
    if (!ckpt_valid) { err = "checkpoint: nothing saved"; return IMC_ERR_STATE; }
    IMC_RC(ensure_capacity(ckpt_host.n_part));
    const CkptHost& c = ckpt_host;
    n_part = c.n_part; n_holes = c.n_holes; alive_known = false; echeck_cached = false; temp_wide = c.temp_wide; red_fixed = c.red_fixed; dep_perm = c.dep_perm;
    fx_mul_dep = c.fx_mul_dep; fx_mul_rad = c.fx_mul_rad; fx_mul_lost = c.fx_mul_lost; rad_total_h = c.rad_total_h; last_mode = c.last_mode;
    totalenergy = c.totalenergy; totalenergydep = c.totalenergydep; radenergyold = c.radenergyold; iterations = c.iterations;
    n_transport_calls = c.n_transport_calls; rate_static = c.rate_static; rate_refill = c.rate_refill; rate_event = c.rate_event;
    hist_n = std::min(hist_n, c.hist_n); hist_dropped = c.hist_dropped;
    last_global_segments = c.last_global_segments; n_global_after_source = c.n_global_after_source;
    auto items = ckpt_items(n_part);   // same order and sizes as at save time: the mesh is fixed, n_part restored above
    size_t off = 0;
    for (auto& it : items) { IMC_CK(cudaMemcpyAsync(it.first, ckpt_blob.p + off, it.second, cudaMemcpyDeviceToDevice, stream)); off += (it.second + 255) & ~(size_t)255; }
    IMC_CK(cudaStreamSynchronize(stream));
    return IMC_OK;
  }
  int get_outcomes(int32_t* ev, int32_t* nseg, int64_t capacity) override {
    IMC_RC(use_device());
    if (out_n == 0) { err = "no outcome record (population above 2^22 or no transport call yet)"; return IMC_ERR_STATE; }
    if (capacity < out_n) { err = "get_outcomes: capacity"; return IMC_ERR_ARG; }
    std::vector<signed char> h(out_n);
    IMC_CK(cudaMemcpyAsync(h.data(), out_event.p, (size_t)out_n, cudaMemcpyDeviceToHost, stream));
    if (nseg) IMC_CK(cudaMemcpyAsync(nseg, out_nseg.p, (size_t)out_n * sizeof(int), cudaMemcpyDeviceToHost, stream));
    IMC_CK(cudaStreamSynchronize(stream));
    if (ev) for (long long i = 0; i < out_n; ++i) ev[i] = h[i];
    return IMC_OK;
  }
};

}  // namespace imc
