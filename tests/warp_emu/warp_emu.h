// warp_emu.h — the warp intrinsics imc_warp_reduce.cuh and imc_warp_runs.cuh use, emulated on the host by 32 threads in lockstep.
//
// TEST INFRASTRUCTURE.  Every lane is a std::thread; an intrinsic is "all lanes publish their operand, wait at a barrier,
// read what they need, wait again".  That is the semantics of the *_sync intrinsics with a full mask, which is the only way
// the code under test calls them (all 32 lanes execute the same sequence of intrinsics).
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#define __device__
#define __forceinline__ inline
#define __restrict__
#define IMC_FULL_MASK 0xffffffffu

namespace warp_emu {
struct Warp {
  std::barrier<> bar{32};
  uint64_t slot[32];
};
inline thread_local Warp* g_warp = nullptr;
inline thread_local int g_lane = 0;
}  // namespace warp_emu

inline unsigned __ballot_sync(unsigned, bool pred) {
  auto* w = warp_emu::g_warp;
  w->slot[warp_emu::g_lane] = pred ? 1u : 0u;
  w->bar.arrive_and_wait();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (unsigned)w->slot[l] << l;
  w->bar.arrive_and_wait();
  return m;
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle operand");
  auto* w = warp_emu::g_warp;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  w->slot[warp_emu::g_lane] = bits;
  w->bar.arrive_and_wait();
  T out;
  std::memcpy(&out, &w->slot[src & 31], sizeof(T));
  w->bar.arrive_and_wait();
  return out;
}
// __shfl_up_sync / __shfl_down_sync: lanes whose source would fall outside the warp keep their own value
template <class T>
inline T __shfl_up_sync(unsigned m, T v, unsigned delta) { const int l = warp_emu::g_lane; return __shfl_sync(m, v, l >= (int)delta ? l - (int)delta : l); }
template <class T>
inline T __shfl_down_sync(unsigned m, T v, unsigned delta) { const int l = warp_emu::g_lane; return __shfl_sync(m, v, l + (int)delta < 32 ? l + (int)delta : l); }
inline int __clz(unsigned m) { return m ? __builtin_clz(m) : 32; }
inline int __ffs(unsigned m) { return m ? __builtin_ctz(m) + 1 : 0; }
using std::signbit;

namespace warp_emu {
// run f(lane) on 32 lockstep lanes; returns lane 0's result
template <class R, class F>
R run_warp(F f) {
  Warp w;
  R res[32];
  std::vector<std::thread> th;
  for (int l = 0; l < 32; ++l)
    th.emplace_back([&, l] { g_warp = &w; g_lane = l; res[l] = f(l); });
  for (auto& t : th) t.join();
  return res[0];
}
}  // namespace warp_emu
