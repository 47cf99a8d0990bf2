// imc_engine.h — host-side interface between the C ABI (imc_capi.cu) and the per-precision CUDA
// engines (imc_engine_f16/f32/f64.cu instantiate EngineT<P> from imc_engine_impl.cuh).
#pragma once
#include <string>
#include <stdint.h>
#include "imc.h"

namespace imc {

struct EngineBase {
  std::string err;
  virtual ~EngineBase() {}
  virtual int init() = 0;
  virtual int set_mesh(const double* dx, const double* dy, const double* sac, const double* sap, const double* ssc,
                       const double* ssp, const double* sstat, const double* bee, const double* rad, const double* temp,
                       const double* tsb, const double* tst, const double* tsl, const double* tsr) = 0;
  virtual int rw_table(double lo, double hi, int n, double* a, double* pr, double* pt) = 0;
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
  virtual void* stream_handle() = 0;
  virtual int history_enable(int64_t) = 0;
  virtual int history_count(int64_t*, int64_t*) = 0;
  virtual int history_get(int, int64_t, int64_t, void*, int64_t) = 0;
  virtual int history_clear() = 0;
  virtual int64_t num_particles() = 0;
  virtual int64_t launches() = 0;
  virtual int get_particles(double*, uint64_t*, int64_t) = 0;
  virtual int set_particles(const double*, const uint64_t*, int64_t) = 0;
  virtual int set_transport_tape(const double*, int, const double*, int, int64_t) = 0;
  virtual int set_source_tape(const double*, int, int64_t) = 0;
  virtual int get_outcomes(int32_t*, int32_t*, int64_t) = 0;
  virtual int sample_planck(int64_t, int64_t, double*) = 0;
  virtual int checkpoint(int) = 0;
};

EngineBase* make_engine_f16(const imc_config& cfg);
EngineBase* make_engine_f32(const imc_config& cfg);
EngineBase* make_engine_f64(const imc_config& cfg);

}  // namespace imc
