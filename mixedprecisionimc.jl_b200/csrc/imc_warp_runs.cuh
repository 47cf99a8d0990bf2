// imc_warp_runs.cuh — runs of equal cells in a particle list, summed by a warp with one deposit per run.
//
// The census tally (imc_tally.jl:84-113) adds E / (dx [dy] scale) of every surviving particle to its cell.  The list is close to
// cell order (particles are emitted cell by cell and the compaction is stable), so consecutive particles mostly share a cell:
// a thread takes a few CONSECUTIVE particles and sums the runs among them in registers (ThreadRuns), then the warp joins the runs
// that continue from lane to lane (warp_join_runs) and each run of the list ends in ONE deposit — an atomic add, which for floats
// in shared memory is a compare-and-swap loop on sm_100 and the limiter of that kernel.
//
// Only *_sync shuffles / ballots with the full mask and __clz: the file also compiles for the host on top of the lockstep
// emulation of tests/warp_emu/warp_emu.h, where tests/test_warp_emu.py checks it against a plain per-cell sum.
#pragma once

namespace imc {

// The runs among the consecutive particles of ONE thread, fed in list order.  `cell_f` / `v_f`: the first run, once a run has
// ended inside the thread (it continues whatever ended in the previous lane); `cell` / `v`: the run still open at the thread's
// last particle.  A run that both starts and ends inside the thread (three or more cells among a thread's particles: rare) is
// deposited at once.  V: double partial sums (float tallies) or exact 64-bit integers (FIXED tallies).
template <class V>
struct ThreadRuns {
  int cell = -1, cell_f = -1;
  V v = V(0), v_f = V(0);
  template <class F>
  __device__ __forceinline__ void push(int cj, V vj, F&& deposit) {
    if (cj == cell) { v += vj; return; }
    if (cell >= 0) {
      if (cell_f < 0) { cell_f = cell; v_f = v; }
      else deposit(cell, v);
    }
    cell = cj; v = vj;
  }
};

// All 32 lanes call this (lanes without particles pass an empty ThreadRuns).  A lane whose particles are one run is a link of a
// longer run; a lane with a run end inside it closes the run of the lanes before it with its FIRST run and opens a new one with
// its LAST run.  head = first lane of a chain of last runs; segmented inclusive scan of the last runs from the head; the lane
// where a chain ends then takes the first run of the next lane if that continues its cell.  On return:
//   want   -> deposit (r.cell, r.v): the run that ends at this lane's last particle, summed over its lanes;
//   want_f -> deposit (r.cell_f, r.v_f): a first run that continues nothing (lane 0, or the previous lane ended in another cell).
template <class V>
__device__ __forceinline__ void warp_join_runs(int lane, ThreadRuns<V>& r, bool& want, bool& want_f) {
  const int prev = __shfl_up_sync(IMC_FULL_MASK, r.cell, 1);
  const bool joins_prev = lane != 0 && prev == (r.cell_f >= 0 ? r.cell_f : r.cell);   // this lane's first particles continue the previous lane's run
  const unsigned heads = __ballot_sync(IMC_FULL_MASK, r.cell_f >= 0 || !joins_prev);  // lane 0 is always a head
  const int head = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
  const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
#pragma unroll
  for (int dlt = 1; dlt < 32; dlt <<= 1) { const V t = __shfl_up_sync(IMC_FULL_MASK, r.v, dlt); if (lane - dlt >= head) r.v += t; }
  const bool give = r.cell_f >= 0 && joins_prev;
  const int take = __shfl_down_sync(IMC_FULL_MASK, give ? 1 : 0, 1);
  const V tv = __shfl_down_sync(IMC_FULL_MASK, r.v_f, 1);
  if (lane != 31 && take) r.v += tv;                   // the next lane is a head, so this lane is the tail of its chain
  want = tail && r.cell >= 0;
  want_f = r.cell_f >= 0 && !joins_prev;
}

}  // namespace imc
