// EngineT<F64> instantiation (one translation unit per precision so the three compile in parallel).
#include "imc_engine_impl.cuh"
namespace imc { EngineBase* make_engine_f64(const imc_config& cfg) { return new EngineT<F64>(cfg); } }
