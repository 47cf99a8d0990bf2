// imc_capi.cu — the C ABI of include/imc.h for the CUDA engine (libimc_b200.so).
#include <new>
#include <string>
#include "imc_engine.h"

using namespace imc;

struct imc_engine { EngineBase* e; };
static thread_local std::string g_create_err;

extern "C" {

int imc_abi_version(void) { return IMC_ABI_VERSION; }
const char* imc_backend(void) { return "cuda-sm_100a"; }

int imc_create(const imc_config* cfg, imc_handle* out) {
  if (!cfg || !out) { g_create_err = "null argument"; return IMC_ERR_ARG; }
  *out = nullptr;
  if (cfg->struct_size != (int32_t)sizeof(imc_config)) { g_create_err = "imc_config size mismatch"; return IMC_ERR_ARG; }
  if ((cfg->geometry != 1 && cfg->geometry != 2) || cfg->nx < 1 || (cfg->geometry == 2 && cfg->ny < 1) ||
      cfg->n_scales < 1 || cfg->n_scales > IMC_MAX_SCALES) { g_create_err = "bad geometry / sizes"; return IMC_ERR_ARG; }
  if (cfg->randomwalk && cfg->geometry != 1) { g_create_err = "RANDOMWALK is 1-D only"; return IMC_ERR_ARG; }
  for (int i = 0; i < (cfg->geometry == 1 ? 2 : 4); ++i)
    if (cfg->bc[i] != IMC_REFLECT && cfg->bc[i] != IMC_VACUUM) { g_create_err = "boundary condition must be REFLECT or VACUUM"; return IMC_ERR_ARG; }
  if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) { g_create_err = "bad rank / world"; return IMC_ERR_ARG; }
  EngineBase* e = nullptr;
  try {
    switch (cfg->precision) {
      case IMC_F16: e = make_engine_f16(*cfg); break;
      case IMC_F32: e = make_engine_f32(*cfg); break;
      case IMC_F64: e = make_engine_f64(*cfg); break;
      default: g_create_err = "bad precision"; return IMC_ERR_ARG;
    }
  } catch (const std::bad_alloc&) { g_create_err = "out of memory"; return IMC_ERR_NOMEM; }
  int rc = e->init();
  if (rc) { g_create_err = e->err; delete e; return rc; }
  *out = new imc_engine{e};
  return IMC_OK;
}
void imc_destroy(imc_handle h) { if (h) { delete h->e; delete h; } }
const char* imc_last_error(imc_handle h) { return h ? h->e->err.c_str() : g_create_err.c_str(); }

#define GUARD(call) do { if (!h) return IMC_ERR_ARG; try { return (call); } catch (const std::bad_alloc&) { h->e->err = "out of host memory"; return IMC_ERR_NOMEM; } } while (0)

int imc_set_mesh(imc_handle h, const double* dx, const double* dy, const double* sac, const double* sap, const double* ssc,
                 const double* ssp, const double* sstat, const double* bee, const double* rad, const double* temp,
                 const double* tsb, const double* tst, const double* tsl, const double* tsr) {
  GUARD(h->e->set_mesh(dx, dy, sac, sap, ssc, ssp, sstat, bee, rad, temp, tsb, tst, tsl, tsr));
}
int imc_rw_table(imc_handle h, double lo, double hi, int32_t n, double* a, double* pr, double* pt) { GUARD(h->e->rw_table(lo, hi, n, a, pr, pt)); }
int imc_update(imc_handle h, double dt) { GUARD(h->e->update(dt)); }
int imc_source(imc_handle h, double dt, int64_t n_input, double cellmin, int64_t step, int64_t ncg, imc_source_stats* out) { GUARD(h->e->source(dt, n_input, cellmin, step, ncg, out)); }
int imc_transport(imc_handle h, double dt, int64_t step, imc_transport_stats* out) { GUARD(h->e->transport(dt, step, out)); }
int imc_clean(imc_handle h, int64_t* n) { GUARD(h->e->clean(n)); }
int imc_tally_local(imc_handle h) { GUARD(h->e->tally_local()); }
int imc_tally_finish(imc_handle h, double t, double dt, imc_tally_stats* out) { GUARD(h->e->tally_finish(t, dt, out)); }
int imc_tally(imc_handle h, double t, double dt, imc_tally_stats* out) {
  if (!h) return IMC_ERR_ARG;
  int rc = imc_tally_local(h);
  return rc ? rc : imc_tally_finish(h, t, dt, out);
}
int imc_energycheck(imc_handle h, imc_energy_stats* out) { GUARD(h->e->energycheck(out)); }
int imc_step(imc_handle h, double t, double dt, int64_t n_input, double cellmin, int64_t step,
             imc_source_stats* src, imc_transport_stats* trk, imc_tally_stats* tal, imc_energy_stats* chk) {
  if (!h) return IMC_ERR_ARG;
  int rc;
  if ((rc = imc_update(h, dt))) return rc;
  if ((rc = imc_source(h, dt, n_input, cellmin, step, -1, src))) return rc;
  if ((rc = imc_transport(h, dt, step, trk))) return rc;
  if ((rc = imc_clean(h, nullptr))) return rc;
  if ((rc = imc_tally(h, t, dt, tal))) return rc;
  return imc_energycheck(h, chk);
}
int imc_reduce_buffer(imc_handle h, void** p, int64_t* n, int32_t* is_int) { GUARD(h->e->reduce_buffer(p, n, is_int)); }
int imc_get_field(imc_handle h, int32_t f, double* dst, int64_t n) { GUARD(h->e->get_field(f, dst, n)); }
int imc_set_state(imc_handle h, const double* t, const double* m, const double* r) { GUARD(h->e->set_state(t, m, r)); }
int32_t imc_field_elsize(imc_handle h, int32_t f) { return h ? h->e->field_elsize(f) : 0; }
int imc_get_field_native(imc_handle h, int32_t f, void* dst, int64_t bytes) { GUARD(h->e->get_field_native(f, dst, bytes)); }
int imc_set_state_native(imc_handle h, const void* t, const void* m, const void* r) { GUARD(h->e->set_state_native(t, m, r)); }
void* imc_stream(imc_handle h) { return h ? h->e->stream_handle() : nullptr; }
int imc_history_enable(imc_handle h, int64_t cap) { GUARD(h->e->history_enable(cap)); }
int imc_history_count(imc_handle h, int64_t* n, int64_t* dropped) { GUARD(h->e->history_count(n, dropped)); }
int imc_history_get(imc_handle h, int32_t f, int64_t first, int64_t count, void* dst, int64_t bytes) { GUARD(h->e->history_get(f, first, count, dst, bytes)); }
int imc_history_clear(imc_handle h) { GUARD(h->e->history_clear()); }
int64_t imc_num_particles(imc_handle h) { return h ? h->e->num_particles() : -1; }
int64_t imc_kernel_launches(imc_handle h) { return h ? h->e->launches() : -1; }
int imc_get_particles(imc_handle h, double* s, uint64_t* ids, int64_t cap) { GUARD(h->e->get_particles(s, ids, cap)); }
int imc_set_particles(imc_handle h, const double* s, const uint64_t* ids, int64_t n) { GUARD(h->e->set_particles(s, ids, n)); }
int imc_set_transport_tape(imc_handle h, const double* u, int32_t nu, const double* e, int32_t ne, int64_t slots) { GUARD(h->e->set_transport_tape(u, nu, e, ne, slots)); }
int imc_set_source_tape(imc_handle h, const double* u, int32_t nu, int64_t slots) { GUARD(h->e->set_source_tape(u, nu, slots)); }
int imc_get_outcomes(imc_handle h, int32_t* ev, int32_t* nseg, int64_t cap) { GUARD(h->e->get_outcomes(ev, nseg, cap)); }
int imc_sample_planck(imc_handle h, int64_t n, int64_t step, double* out) { GUARD(h->e->sample_planck(n, step, out)); }
int imc_checkpoint(imc_handle h, int32_t op) { GUARD(h->e->checkpoint(op)); }

}  // extern "C"
