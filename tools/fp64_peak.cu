// fp64_peak.cu -- measures the fp64 roofs the moment deposition is judged against, on the GPU it runs on:
//   DFMA issue rate (vector fp64 pipe), DMMA m8n8k4 issue rate (fp64 tensor path), both mixed in one
//   instruction stream (do they share a pipe?), and the wavefront cost of the 128-bit shared-memory load
//   patterns of k_cell_moments (partial broadcast).  CUDA events around each kernel, clock64() inside;
//   prints ONE JSON object.  bench.py runs it (xpic_b200/_build/fp64_peak) and reads "dmma_tflops" as the
//   peak of the tensor roofline; profiles/ keeps a copy of the output.
//     nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_));                  \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// MODE 0: 8 DFMA chains; 1: 8 DMMA chains; 2: 8 DMMA + 8 DFMA per iteration
template <int MODE>
__global__ void __launch_bounds__(256) k_fp64(int iters, double seed, double* out, long long* cycles)
{
  double acc[8][2], f[8];
  const double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-9 * seed;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    acc[k][0] = acc[k][1] = 0.0;
    f[k] = k * 0.5;
  }
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE != 0) dmma(acc[k][0], acc[k][1], a, b);
      if (MODE != 1) f[k] = fma(f[k], b, a);
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += acc[k][0] + acc[k][1] + f[k];
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// shared-memory probe: every lane issues `iters` x 8 loads of WIDTH bytes with the address pattern PAT
//   PAT 0: all lanes distinct, consecutive (the bandwidth floor: WIDTH * 32 / 128 wavefronts)
//   PAT 1: the weight load of k_cell_moments: lane (gq, q) reads the 16-byte chunk (4 * (gq & 1)) of record q
//          (records 13 chunks apart): 8 distinct chunks in 8 different bank groups
//   PAT 2: the alpha load: chunk 3 of record q: 4 distinct chunks
template <int PAT>
__global__ void __launch_bounds__(256) k_lds(int iters, double* out, long long* cycles)
{
  __shared__ __align__(16) double sm[8 * 480];
  for (int i = threadIdx.x; i < 8 * 480; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, gq = lane >> 2, q = lane & 3;
  const double* base = sm + wid * 480;
  int ofs;
  if (PAT == 0)
    ofs = 2 * lane;
  else if (PAT == 1)
    ofs = q * 26 + 2 * (4 * (gq & 1));
  else
    ofs = q * 26 + 6;
  double2 s = make_double2(0.0, 0.0);
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double2 v = *reinterpret_cast<const double2*>(base + ofs + ((i + k) & 3) * 104);
      s.x += v.x;
      s.y += v.y;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s.x + s.y;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <class K>
static int time_kernel(K launch, float* ms)
{
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch();  // warm-up
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float t;
    CK(cudaEventElapsedTime(&t, e0, e1));
    if (t < best) best = t;
  }
  *ms = best;
  return 0;
}

int main()
{
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 8, threads = 256, iters = 20000;
  double* out;
  long long* cyc;
  CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  CK(cudaMalloc(&cyc, sizeof(long long)));
  float ms[3];
  long long cycles[3];
  for (int mode = 0; mode < 3; ++mode) {
    auto launch = [&]() {
      if (mode == 0) k_fp64<0><<<blocks, threads>>>(iters, 1.0, out, cyc);
      if (mode == 1) k_fp64<1><<<blocks, threads>>>(iters, 1.0, out, cyc);
      if (mode == 2) k_fp64<2><<<blocks, threads>>>(iters, 1.0, out, cyc);
    };
    if (time_kernel(launch, &ms[mode])) return 1;
    CK(cudaMemcpy(&cycles[mode], cyc, sizeof(long long), cudaMemcpyDeviceToHost));
  }
  const double nthreads = (double)blocks * threads, nwarps = nthreads / 32;
  const double dfma_tf = 2.0 * nthreads * iters * 8 / (ms[0] * 1e-3) / 1e12;
  const double dmma_tf = 512.0 * nwarps * iters * 8 / (ms[1] * 1e-3) / 1e12;
  const double mix_dmma_tf = 512.0 * nwarps * iters * 8 / (ms[2] * 1e-3) / 1e12;
  const double mix_dfma_tf = 2.0 * nthreads * iters * 8 / (ms[2] * 1e-3) / 1e12;
  // 16 warps per SM = 4 per sub-partition: cycles one sub-partition spends per instruction
  const double warps_per_smsp = 8.0 * threads / 32 / 4;
  const double dfma_cyc = cycles[0] / (iters * 8.0 * warps_per_smsp);
  const double dmma_cyc = cycles[1] / (iters * 8.0 * warps_per_smsp);
  const double mix_cyc = cycles[2] / (iters * 8.0 * warps_per_smsp);

  float lms[3];
  long long lcyc[3];
  const int liters = 4000;
  for (int pat = 0; pat < 3; ++pat) {
    auto launch = [&]() {
      if (pat == 0) k_lds<0><<<sms, 256>>>(liters, out, cyc);
      if (pat == 1) k_lds<1><<<sms, 256>>>(liters, out, cyc);
      if (pat == 2) k_lds<2><<<sms, 256>>>(liters, out, cyc);
    };
    if (time_kernel(launch, &lms[pat])) return 1;
    CK(cudaMemcpy(&lcyc[pat], cyc, sizeof(long long), cudaMemcpyDeviceToHost));
  }
  // one CTA of 8 warps per SM: cycles the SM's shared-memory pipe spends per warp-wide 128-bit load
  double lds_cyc[3];
  for (int pat = 0; pat < 3; ++pat) lds_cyc[pat] = lcyc[pat] / (liters * 8.0 * 8.0);

  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_max\": %.0f, "
         "\"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f, \"mixed_dmma_tflops\": %.3f, \"mixed_dfma_tflops\": %.3f, "
         "\"cycles_per_dfma_per_smsp\": %.3f, \"cycles_per_dmma_per_smsp\": %.3f, \"cycles_per_dmma_plus_dfma_per_smsp\": %.3f, "
         "\"ms\": [%.4f, %.4f, %.4f], "
         "\"lds128_cycles_per_warp_load\": {\"all_distinct\": %.3f, \"weights_8_chunks\": %.3f, \"alpha_4_chunks\": %.3f}, "
         "\"how\": \"%d CTAs x %d threads, %d iterations x 8 independent chains per thread; best of 5 launches, CUDA events; "
         "cycles from clock64() of one thread\"}\n",
         prop.name, sms, khz / 1e3, dfma_tf, dmma_tf, mix_dmma_tf, mix_dfma_tf, dfma_cyc, dmma_cyc, mix_cyc, ms[0], ms[1], ms[2], lds_cyc[0], lds_cyc[1],
         lds_cyc[2], blocks, threads, iters);
  cudaFree(out);
  cudaFree(cyc);
  return 0;
}
