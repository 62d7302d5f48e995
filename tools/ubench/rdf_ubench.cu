// Micro-benchmarks that drive the K3 (all-pairs RDF) kernel design on B200.
//   part A: shared-memory atomic (ATOMS) throughput under histogram-like access patterns
//   part B: K3 prototype (fixed-point min-image, register i-tile x smem j-broadcast) with and
//           without the shared histogram, to separate FP32-issue cost from shared-atomic cost.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/rdf_ubench tools/ubench/rdf_ubench.cu
// Not part of the product; results are summarised in profiles/.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// ---------------------------------------------------------------- part A
template <int MODE>
__global__ void __launch_bounds__(256) atoms_bench(int iters, unsigned nbins, unsigned* gout, long long* cyc,
                                                   unsigned* ghist)
{
  extern __shared__ unsigned hist[];
  for (unsigned i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  unsigned acc = 0;
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp_base = ((threadIdx.x >> 5) * 397u) % (nbins - 320u);
  long long t0 = clock64();
#pragma unroll 8
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    unsigned bin = __umulhi(s, nbins);
    if (MODE == 0) { acc ^= bin; }
    if (MODE == 1) { atomicAdd(&hist[bin], 1u); }
    if (MODE == 2) { if (s & 0x10000u) atomicAdd(&hist[bin], 1u); }
    if (MODE == 3) { volatile unsigned* h = hist; h[bin] = h[bin] + 1u; }
    if (MODE == 4) { unsigned b = warp_base + (__umulhi(s, 300u)); atomicAdd(&hist[b], 1u); }
    if (MODE == 5) { unsigned b = warp_base + lane + (it & 63); atomicAdd(&hist[b], 1u); }  // conflict free
    if (MODE == 6) { atomicAdd(&ghist[(size_t)blockIdx.x * nbins + bin], 1u); }
    if (MODE == 7) { unsigned m = __match_any_sync(0xffffffffu, bin); if ((__ffs(m) - 1) == (int)lane) atomicAdd(&hist[bin], (unsigned)__popc(m)); }
    if (MODE == 8) { if ((s & 0x30000u) == 0) atomicAdd(&hist[bin], 1u); }  // 25% lanes
  }
  long long t1 = clock64();
  __syncthreads();
  for (unsigned i = threadIdx.x; i < nbins; i += blockDim.x) acc += hist[i];
  if (acc == 0xdeadbeefu) gout[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run_atoms(const char* name, int nsm, unsigned nbins)
{
  const int blocks_per_sm = 4, threads = 256, iters = 4096;
  int grid = nsm * blocks_per_sm;
  unsigned* gout; long long* cyc; unsigned* ghist;
  CK(cudaMalloc(&gout, 4)); CK(cudaMalloc(&cyc, grid * sizeof(long long)));
  CK(cudaMalloc(&ghist, (size_t)grid * nbins * 4)); CK(cudaMemset(ghist, 0, (size_t)grid * nbins * 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  size_t smem = nbins * 4;
  atoms_bench<MODE><<<grid, threads, smem>>>(iters, nbins, gout, cyc, ghist);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  atoms_bench<MODE><<<grid, threads, smem>>>(iters, nbins, gout, cyc, ghist);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h(grid); CK(cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean = 0; for (auto v : h) mean += (double)v; mean /= grid;
  double warp_instr_per_sm = (double)iters * blocks_per_sm * (threads / 32);
  printf("A %-34s nbins=%5u  %8.3f ms  cyc/blk=%10.0f  cyc per warp-op per SM = %7.3f   lane-ops/s = %.3e\n",
         name, nbins, ms, mean, mean / warp_instr_per_sm, (double)grid * threads * iters / (ms * 1e-3));
  CK(cudaFree(gout)); CK(cudaFree(cyc)); CK(cudaFree(ghist));
}

// ---------------------------------------------------------------- part B
// coordinates as 32-bit box fractions: x_fix = rint(x/L * 2^32) mod 2^32; integer difference wraps = minimum image.
#define MAGIC_BITS 0x4B000000u
template <int HIST, int IPT, int THREADS, int FLOATPATH>
__global__ void __launch_bounds__(THREADS) rdf_proto(const int4* __restrict__ A, const int4* __restrict__ B,
                                                     int n0, int n1, int jchunk, unsigned rbin, float scale_lo,
                                                     float scale_hi, float invL_f, float L_f,
                                                     unsigned long long* ghist, unsigned long long* gamb)
{
  extern __shared__ unsigned smem[];
  unsigned* hist = smem;                       // rbin u32
  int4* sj = (int4*)(smem + ((rbin + 3) & ~3u));  // TJ entries
  const int TJ = 1024;
  for (unsigned i = threadIdx.x; i < rbin; i += THREADS) hist[i] = 0;
  const int ibase = blockIdx.x * (THREADS * IPT);
  int xi[IPT], yi[IPT], zi[IPT];
#pragma unroll
  for (int q = 0; q < IPT; ++q) {
    int i = ibase + threadIdx.x + q * THREADS;
    int4 p = (i < n0) ? A[i] : make_int4(0, 0, 0, 0);
    xi[q] = p.x; yi[q] = p.y; zi[q] = p.z;
  }
  const int j0 = blockIdx.y * jchunk;
  const int j1 = min(n1, j0 + jchunk);
  unsigned namb = 0;
  unsigned acc = 0;
  const unsigned hist_cbase = (unsigned)__cvta_generic_to_shared(hist) - MAGIC_BITS * 4u;
  const unsigned dummy_addr = (unsigned)__cvta_generic_to_shared(sj + TJ) + (threadIdx.x & 31) * 4u;
  unsigned one; asm volatile("mov.u32 %0, 1;" : "=r"(one));
  for (int jt = j0; jt < j1; jt += TJ) {
    __syncthreads();
    int cnt = min(TJ, j1 - jt);
    for (int t = threadIdx.x; t < cnt; t += THREADS) sj[t] = B[jt + t];
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < cnt; ++j) {
      int4 pj = sj[j];
#pragma unroll
      for (int q = 0; q < IPT; ++q) {
        float r2;
        if (FLOATPATH) {
          float fx = __int_as_float(xi[q]) - __int_as_float(pj.x);
          float fy = __int_as_float(yi[q]) - __int_as_float(pj.y);
          float fz = __int_as_float(zi[q]) - __int_as_float(pj.z);
          fx = fmaf(-L_f, rintf(fx * invL_f), fx);
          fy = fmaf(-L_f, rintf(fy * invL_f), fy);
          fz = fmaf(-L_f, rintf(fz * invL_f), fz);
          r2 = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
        } else {
          float fx = (float)(xi[q] - pj.x);
          float fy = (float)(yi[q] - pj.y);
          float fz = (float)(zi[q] - pj.z);
          r2 = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
        }
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
        unsigned k1 = __float_as_uint(fmaf(r, scale_lo, 8388607.5f));
        unsigned k2 = __float_as_uint(fmaf(r, scale_hi, 8388607.5f));
        bool in = (k1 - (MAGIC_BITS + 32u)) < (rbin - 32u);
        bool cnt_ok = in && (k1 == k2);
        bool amb = (!cnt_ok) && (k1 < MAGIC_BITS + rbin);
        if (HIST == 0) { acc += cnt_ok ? k1 : 0u; }
        if (HIST == 1) {
          unsigned addr = k1 * 4u + hist_cbase;
          asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 0; @p red.shared.add.u32 [%0], 1; }" :: "r"(addr), "r"((unsigned)cnt_ok) : "memory");
        }
        if (HIST == 2) {  // branch-free: losers go to a per-lane dummy word behind the histogram
          unsigned addr = cnt_ok ? (k1 * 4u + hist_cbase) : dummy_addr;
          asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(addr) : "memory");
        }
        if (HIST == 3) {  // ATOMS.ADD with an opaque register operand (0 for losers)
          unsigned addr = cnt_ok ? (k1 * 4u + hist_cbase) : dummy_addr;
          asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(one) : "memory");
        }
        namb += amb ? 1u : 0u;
      }
    }
  }
  __syncthreads();
  if (HIST == 1) {
    for (unsigned i = threadIdx.x; i < rbin; i += THREADS)
      if (hist[i]) atomicAdd(&ghist[i], (unsigned long long)hist[i]);
  } else {
    if (acc == 0xdeadbeefu) ghist[0] = acc;
  }
  // block-level ambiguous count
  for (int o = 16; o > 0; o >>= 1) namb += __shfl_xor_sync(0xffffffffu, namb, o);
  if ((threadIdx.x & 31) == 0 && namb) atomicAdd(gamb, (unsigned long long)namb);
}


// ---- v2: unconditional ATOMS into an oversize histogram (all reachable k), trash word at hist[-1],
//      count-then-correct ambiguity handling (only detection cost modelled here), IMAD addressing.
template <int IPT, int THREADS, int VAR>
__global__ void __launch_bounds__(THREADS) rdf_proto2(const int4* __restrict__ A, const int4* __restrict__ B,
                                                      int n0, int n1, int jchunk, unsigned rbin, unsigned kmax,
                                                      float scale_lo, float scale_hi,
                                                      unsigned long long* ghist, unsigned long long* gamb)
{
  extern __shared__ unsigned smem[];
  const int TJ = 1024;
  int4* sj = (int4*)smem;                    // TJ entries
  unsigned* hist = smem + TJ * 4 + 4;        // hist[-1] is a trash word
  for (unsigned i = threadIdx.x; i < kmax + 4; i += THREADS) hist[(int)i - 4] = 0;
  const int ibase = blockIdx.x * (THREADS * IPT);
  int xi[IPT], yi[IPT], zi[IPT];
#pragma unroll
  for (int q = 0; q < IPT; ++q) {
    int i = ibase + threadIdx.x + q * THREADS;
    int4 p = (i < n0) ? A[i] : make_int4(0, 0, 0, 0);
    xi[q] = p.x; yi[q] = p.y; zi[q] = p.z;
  }
  const int j0 = blockIdx.y * jchunk;
  const int j1 = min(n1, j0 + jchunk);
  unsigned namb = 0;
  const unsigned hist_cbase = (unsigned)__cvta_generic_to_shared(hist) - MAGIC_BITS * 4u;
  for (int jt = j0; jt < j1; jt += TJ) {
    __syncthreads();
    int cnt = min(TJ, j1 - jt);
    for (int t = threadIdx.x; t < cnt; t += THREADS) sj[t] = B[jt + t];
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < cnt; ++j) {
      int4 pj = sj[j];
      unsigned dacc = 0, kmin = 0xffffffffu;
#pragma unroll
      for (int q = 0; q < IPT; ++q) {
        int dx, dy, dz;
        if (VAR & 1) {  // force IMAD-pipe subtraction
          asm("mad.lo.s32 %0, %1, -1, %2;" : "=r"(dx) : "r"(pj.x), "r"(xi[q]));
          asm("mad.lo.s32 %0, %1, -1, %2;" : "=r"(dy) : "r"(pj.y), "r"(yi[q]));
          asm("mad.lo.s32 %0, %1, -1, %2;" : "=r"(dz) : "r"(pj.z), "r"(zi[q]));
        } else { dx = xi[q] - pj.x; dy = yi[q] - pj.y; dz = zi[q] - pj.z; }
        float fx = (float)dx, fy = (float)dy, fz = (float)dz;
        float r2 = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
        unsigned k1 = __float_as_uint(fmaf(r - 2.0f, scale_lo, 8388607.5f));
        unsigned k2 = __float_as_uint(fmaf(r + 2.0f, scale_hi, 8388607.5f));
        unsigned addr;
        asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(addr) : "r"(k1), "r"(hist_cbase));
        asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(addr) : "memory");
        if (VAR & 2) { dacc += k2 - k1; kmin = min(kmin, k1); }
        else { namb += ((k1 != k2) || (k1 <= MAGIC_BITS)) ? 1u : 0u; }
      }
      if (VAR & 2) { if (dacc != 0 || kmin <= MAGIC_BITS) namb++; }
    }
  }
  __syncthreads();
  for (unsigned i = threadIdx.x; i < rbin; i += THREADS)
    if (hist[i]) atomicAdd(&ghist[i], (unsigned long long)hist[i]);
  for (int o = 16; o > 0; o >>= 1) namb += __shfl_xor_sync(0xffffffffu, namb, o);
  if ((threadIdx.x & 31) == 0 && namb) atomicAdd(gamb, (unsigned long long)namb);
}

static uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }

template <int HIST, int IPT, int THREADS, int FLOATPATH>
static void run_proto(const char* name, int N, int jsplit, int reps)
{
  const double L = 12.913; const double Rm = L / 2; const unsigned rbin = (unsigned)(Rm / 0.002);
  std::vector<int4> h(N);
  uint64_t s = 20200102;
  for (int i = 0; i < N; ++i) {
    if (FLOATPATH) {
      float x = (float)((splitmix(s) >> 40) * (1.0 / 16777216.0) * L), y = (float)((splitmix(s) >> 40) * (1.0 / 16777216.0) * L), z = (float)((splitmix(s) >> 40) * (1.0 / 16777216.0) * L);
      h[i] = make_int4(*(int*)&x, *(int*)&y, *(int*)&z, 0);
    } else {
      h[i] = make_int4((int)(uint32_t)splitmix(s), (int)(uint32_t)splitmix(s), (int)(uint32_t)splitmix(s), 0);
    }
  }
  int4* d; CK(cudaMalloc(&d, N * sizeof(int4))); CK(cudaMemcpy(d, h.data(), N * sizeof(int4), cudaMemcpyHostToDevice));
  unsigned long long* ghist; unsigned long long* gamb;
  CK(cudaMalloc(&ghist, rbin * 8)); CK(cudaMalloc(&gamb, 8));
  CK(cudaMemset(ghist, 0, rbin * 8)); CK(cudaMemset(gamb, 0, 8));
  double unit = FLOATPATH ? 1.0 : L / 4294967296.0;
  double scale = unit * rbin / Rm;
  float scale_lo = (float)(scale * (1 - 4e-7)), scale_hi = (float)(scale * (1 + 4e-7));
  int itiles = (N + THREADS * IPT - 1) / (THREADS * IPT);
  int jchunk = (N + jsplit - 1) / jsplit;
  dim3 grid(itiles, jsplit);
  size_t smem = ((rbin + 3) & ~3u) * 4 + 1024 * sizeof(int4) + 128;
  CK(cudaFuncSetAttribute(rdf_proto<HIST, IPT, THREADS, FLOATPATH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rdf_proto<HIST, IPT, THREADS, FLOATPATH>, THREADS, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  rdf_proto<HIST, IPT, THREADS, FLOATPATH><<<grid, THREADS, smem>>>(d, d, N, N, jchunk, rbin, scale_lo, scale_hi, (float)(1.0 / L), (float)L, ghist, gamb);
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(ghist, 0, rbin * 8)); CK(cudaMemset(gamb, 0, 8));
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r)
    rdf_proto<HIST, IPT, THREADS, FLOATPATH><<<grid, THREADS, smem>>>(d, d, N, N, jchunk, rbin, scale_lo, scale_hi, (float)(1.0 / L), (float)L, ghist, gamb);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  std::vector<unsigned long long> hh(rbin); CK(cudaMemcpy(hh.data(), ghist, rbin * 8, cudaMemcpyDeviceToHost));
  unsigned long long amb; CK(cudaMemcpy(&amb, gamb, 8, cudaMemcpyDeviceToHost));
  unsigned long long tot = 0; for (auto v : hh) tot += v;
  double pairs = (double)N * N;
  printf("B %-40s N=%d grid=(%d,%d) occ=%d  %8.3f ms/frame  %.3e pairs/s  counted=%.4f amb=%.5f%%\n", name, N, itiles, jsplit, occ, ms,
         pairs / (ms * 1e-3), (double)tot / reps / pairs, 100.0 * (double)amb / reps / pairs);
  CK(cudaFree(d)); CK(cudaFree(ghist)); CK(cudaFree(gamb));
}

template <int IPT, int THREADS, int VAR>
static void run_proto2(const char* name, int N, int jsplit, int reps)
{
  const double L = 12.913; const double Rm = L / 2; const unsigned rbin = (unsigned)(Rm / 0.002);
  std::vector<int4> h(N);
  uint64_t s = 20200102;
  for (int i = 0; i < N; ++i) h[i] = make_int4((int)(uint32_t)splitmix(s), (int)(uint32_t)splitmix(s), (int)(uint32_t)splitmix(s), 0);
  int4* d; CK(cudaMalloc(&d, N * sizeof(int4))); CK(cudaMemcpy(d, h.data(), N * sizeof(int4), cudaMemcpyHostToDevice));
  unsigned long long* ghist; unsigned long long* gamb;
  CK(cudaMalloc(&ghist, rbin * 8)); CK(cudaMalloc(&gamb, 8));
  CK(cudaMemset(ghist, 0, rbin * 8)); CK(cudaMemset(gamb, 0, 8));
  double unit = L / 4294967296.0;
  double scale = unit * rbin / Rm;
  unsigned kmax = (unsigned)(std::sqrt(3.0) * 0.5 * L / unit * scale) + 8;
  float scale_lo = (float)(scale * (1 - 4e-7)), scale_hi = (float)(scale * (1 + 4e-7));
  int itiles = (N + THREADS * IPT - 1) / (THREADS * IPT);
  int jchunk = (N + jsplit - 1) / jsplit;
  dim3 grid(itiles, jsplit);
  size_t smem = 1024 * sizeof(int4) + 16 + (kmax + 4) * 4;
  auto kern = rdf_proto2<IPT, THREADS, VAR>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<<<grid, THREADS, smem>>>(d, d, N, N, jchunk, rbin, kmax, scale_lo, scale_hi, ghist, gamb);
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(ghist, 0, rbin * 8)); CK(cudaMemset(gamb, 0, 8));
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r)
    kern<<<grid, THREADS, smem>>>(d, d, N, N, jchunk, rbin, kmax, scale_lo, scale_hi, ghist, gamb);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  std::vector<unsigned long long> hh(rbin); CK(cudaMemcpy(hh.data(), ghist, rbin * 8, cudaMemcpyDeviceToHost));
  unsigned long long amb; CK(cudaMemcpy(&amb, gamb, 8, cudaMemcpyDeviceToHost));
  unsigned long long tot = 0; for (auto v : hh) tot += v;
  double pairs = (double)N * N;
  printf("C %-40s N=%d grid=(%d,%d) occ=%d kmax=%u %8.3f ms/frame  %.3e pairs/s  counted=%.4f amb=%.5f%%\n", name, N, itiles, jsplit, occ, kmax, ms,
         pairs / (ms * 1e-3), (double)tot / reps / pairs, 100.0 * (double)amb / reps / pairs);
  CK(cudaFree(d)); CK(cudaFree(ghist)); CK(cudaFree(gamb));
}

int main(int argc, char** argv)
{
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf("device %s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, clk);
  int nsm = p.multiProcessorCount;
  run_atoms<0>("baseline (no smem op)", nsm, 3228);
  run_atoms<1>("ATOMS uniform random, 32 lanes", nsm, 3228);
  run_atoms<2>("ATOMS uniform random, ~50% lanes", nsm, 3228);
  run_atoms<8>("ATOMS uniform random, ~25% lanes", nsm, 3228);
  run_atoms<3>("LDS+STS (racy RMW) random", nsm, 3228);
  run_atoms<4>("ATOMS clustered 300-bin window", nsm, 3228);
  run_atoms<5>("ATOMS conflict-free consecutive", nsm, 3228);
  run_atoms<6>("RED.global per-block hist random", nsm, 3228);
  run_atoms<7>("match_any + leader ATOMS", nsm, 3228);
  run_atoms<1>("ATOMS uniform random, 32 lanes", nsm, 16140);
  int N = 72000;
  if (argc > 1) N = atoi(argv[1]);
  if (argc > 2) {  // short mode for ncu
    run_proto<2, 4, 128, 0>("fixed  IPT4 T128 ATOMS redirect", N, 8, 1);
    run_proto2<4, 128, 0>("v2 IPT4 T128", N, 8, 1);
    run_proto2<4, 128, 3>("v2 IPT4 T128 imadsub+acc", N, 8, 1);
    return 0;
  }
  run_proto2<4, 128, 0>("v2 IPT4 T128", N, 8, 3);
  run_proto2<4, 128, 1>("v2 IPT4 T128 imadsub", N, 8, 3);
  run_proto2<4, 128, 2>("v2 IPT4 T128 acc", N, 8, 3);
  run_proto2<4, 128, 3>("v2 IPT4 T128 imadsub+acc", N, 8, 3);
  run_proto2<4, 256, 0>("v2 IPT4 T256", N, 16, 3);
  run_proto2<4, 256, 3>("v2 IPT4 T256 imadsub+acc", N, 16, 3);
  run_proto2<2, 256, 0>("v2 IPT2 T256", N, 8, 3);
  run_proto2<8, 128, 0>("v2 IPT8 T128", N, 16, 3);
  run_proto2<6, 128, 3>("v2 IPT6 T128 imadsub+acc", N, 12, 3);
  run_proto<0, 4, 128, 0>("fixed  IPT4 T128 no-hist", N, 8, 3);
  run_proto<1, 4, 128, 0>("fixed  IPT4 T128 ATOMS", N, 8, 3);
  run_proto<2, 4, 128, 0>("fixed  IPT4 T128 ATOMS redirect", N, 8, 3);
  run_proto<3, 4, 128, 0>("fixed  IPT4 T128 ATOMS.ADD reg redirect", N, 8, 3);
  run_proto<0, 4, 256, 0>("fixed  IPT4 T256 no-hist", N, 16, 3);
  run_proto<1, 4, 256, 0>("fixed  IPT4 T256 ATOMS", N, 16, 3);
  run_proto<0, 2, 256, 0>("fixed  IPT2 T256 no-hist", N, 8, 3);
  run_proto<1, 2, 256, 0>("fixed  IPT2 T256 ATOMS", N, 8, 3);
  run_proto<0, 8, 128, 0>("fixed  IPT8 T128 no-hist", N, 16, 3);
  run_proto<1, 8, 128, 0>("fixed  IPT8 T128 ATOMS", N, 16, 3);
  run_proto<0, 4, 128, 1>("float  IPT4 T128 no-hist", N, 8, 3);
  run_proto<1, 4, 128, 1>("float  IPT4 T128 ATOMS", N, 8, 3);
  return 0;
}
