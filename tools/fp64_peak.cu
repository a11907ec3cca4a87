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

// shared-memory probe: every lane issues `iters` x 8 loads of W bytes (4, 8 or 16) with the address pattern PAT
//   PAT 0: all lanes distinct, consecutive
//   PAT 1: the weight load of k_cell_moments: lane (gq, q) reads a chunk of record q that depends on a bit of gq
//          (records 13 x 16 bytes apart): 8 distinct addresses
//   PAT 2: the alpha load: one chunk of record q: 4 distinct addresses (8 lanes share each)
// Result: cycles of the SM's shared-memory pipe per warp-wide load (one CTA of 8 warps per SM keeps it busy).
template <int W, int PAT>
__global__ void __launch_bounds__(256) k_lds(int iters, double* out)
{
  __shared__ __align__(16) double sm[8 * 480];
  for (int i = threadIdx.x; i < 8 * 480; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, gq = lane >> 2, q = lane & 3;
  const char* base = reinterpret_cast<const char*>(sm + wid * 480);
  int ofs;  // bytes
  if (PAT == 0)
    ofs = W * lane;
  else if (PAT == 1)
    ofs = q * 208 + 64 * (gq & 1);
  else
    ofs = q * 208 + 48;
  // integer accumulators, one per load of the unrolled body: no dependent fp64 chain limits the issue rate
  unsigned acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const char* p = base + ofs + ((i + k) & 3) * 832;
      if (W == 16) {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        acc[k] += v.x ^ v.y ^ v.z ^ v.w;
      }
      else if (W == 8) {
        const uint2 v = *reinterpret_cast<const uint2*>(p);
        acc[k] += v.x ^ v.y;
      }
      else {
        acc[k] += *reinterpret_cast<const unsigned*>(p);
      }
    }
  }
  unsigned t = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) t ^= acc[k];
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = (double)t;
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
  const double hz = khz * 1e3;  // cycles from the event time at the maximum SM clock (the clocks line of bench.py shows the GPU holds it)
  const double dfma_cyc = ms[0] * 1e-3 * hz / (iters * 8.0 * warps_per_smsp);
  const double dmma_cyc = ms[1] * 1e-3 * hz / (iters * 8.0 * warps_per_smsp);
  const double mix_cyc = ms[2] * 1e-3 * hz / (iters * 8.0 * warps_per_smsp);
  (void)cycles;

  // nine probes: widths 4 / 8 / 16 bytes x patterns 0 / 1 / 2
  float lms[9];
  const int liters = 4000;
  for (int t = 0; t < 9; ++t) {
    auto launch = [&]() {
      switch (t) {
        case 0: k_lds<4, 0><<<sms, 256>>>(liters, out); break;
        case 1: k_lds<4, 1><<<sms, 256>>>(liters, out); break;
        case 2: k_lds<4, 2><<<sms, 256>>>(liters, out); break;
        case 3: k_lds<8, 0><<<sms, 256>>>(liters, out); break;
        case 4: k_lds<8, 1><<<sms, 256>>>(liters, out); break;
        case 5: k_lds<8, 2><<<sms, 256>>>(liters, out); break;
        case 6: k_lds<16, 0><<<sms, 256>>>(liters, out); break;
        case 7: k_lds<16, 1><<<sms, 256>>>(liters, out); break;
        default: k_lds<16, 2><<<sms, 256>>>(liters, out); break;
      }
    };
    if (time_kernel(launch, &lms[t])) return 1;
  }
  // one CTA of 8 warps per SM: SM cycles (at the maximum clock) per warp-wide load
  double lds_cyc[9];
  for (int t = 0; t < 9; ++t) lds_cyc[t] = lms[t] * 1e-3 * khz * 1e3 / (liters * 8.0 * 8.0);

  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_max\": %.0f, "
         "\"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f, \"mixed_dmma_tflops\": %.3f, \"mixed_dfma_tflops\": %.3f, "
         "\"cycles_per_dfma_per_smsp\": %.3f, \"cycles_per_dmma_per_smsp\": %.3f, \"cycles_per_dmma_plus_dfma_per_smsp\": %.3f, "
         "\"ms\": [%.4f, %.4f, %.4f], "
         "\"lds_cycles_per_warp_load\": {\"b32\": [%.3f, %.3f, %.3f], \"b64\": [%.3f, %.3f, %.3f], \"b128\": [%.3f, %.3f, %.3f], "
         "\"patterns\": \"all lanes distinct | 8 distinct addresses | 4 distinct addresses (8 lanes each)\"}, "
         "\"how\": \"%d CTAs x %d threads, %d iterations x 8 independent chains per thread; best of 5 launches, CUDA events; "
         "cycles = event time x maximum SM clock\"}\n",
         prop.name, sms, khz / 1e3, dfma_tf, dmma_tf, mix_dmma_tf, mix_dfma_tf, dfma_cyc, dmma_cyc, mix_cyc, ms[0], ms[1], ms[2], lds_cyc[0], lds_cyc[1],
         lds_cyc[2], lds_cyc[3], lds_cyc[4], lds_cyc[5], lds_cyc[6], lds_cyc[7], lds_cyc[8], blocks, threads, iters);
  cudaFree(out);
  cudaFree(cyc);
  return 0;
}
