// micro-benchmark: how fast can a B200 gather 16-byte bound pairs through random 32-bit indices?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int MODE, int U>
__global__ void gather(const int* __restrict__ idx, const double2* __restrict__ tab, double* out, long long n)
{
   const long long stride = (long long)gridDim.x * blockDim.x * U;
   double acc = 0.0;
   for( long long i = ((long long)blockIdx.x * blockDim.x) * U + threadIdx.x; i < n; i += stride )
   {
      int j[U];
#pragma unroll
      for( int k = 0; k < U; ++k )
         j[k] = (i + (long long)k * blockDim.x < n) ? __ldcs(idx + i + (long long)k * blockDim.x) : 0;
      double2 v[U];
#pragma unroll
      for( int k = 0; k < U; ++k )
      {
         if( MODE == 0 ) v[k] = tab[j[k]];
         else if( MODE == 1 ) v[k] = __ldg(tab + j[k]);
         else if( MODE == 2 ) v[k] = __ldcg(tab + j[k]);
         else { double x = __ldg(reinterpret_cast<const double*>(tab + j[k])); v[k] = make_double2(x, x); }
      }
#pragma unroll
      for( int k = 0; k < U; ++k )
         acc += v[k].x + v[k].y;
   }
   if( acc == 1.2345 )
      out[0] = acc;
}

template <int MODE, int U>
float run(const int* idx, const double2* tab, double* out, long long n, int blocks, int threads)
{
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   for( int w = 0; w < 3; ++w )
      gather<MODE, U><<<blocks, threads>>>(idx, tab, out, n);
   cudaEventRecord(e0);
   const int reps = 10;
   for( int r = 0; r < reps; ++r )
      gather<MODE, U><<<blocks, threads>>>(idx, tab, out, n);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   return ms / reps * 1000.f;
}

int main()
{
   const long long n = 10000000;
   const int ncols = 1000000;
   std::vector<int> h(n);
   srand(1);
   for( long long i = 0; i < n; ++i )
      h[i] = (int)(((long long)rand() * 32768 + rand()) % ncols);
   int* idx; double2* tab; double* out;
   cudaMalloc(&idx, n * 4); cudaMalloc(&tab, (size_t)ncols * 16); cudaMalloc(&out, 8);
   cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
   cudaMemset(tab, 0, (size_t)ncols * 16);
   printf("10M random gathers of 16 B from a 16 MB table (+ 40 MB index stream); us per pass\n");
   const int occs[] = {2, 4, 8};
   for( int oi = 0; oi < 3; ++oi )
   {
      const int blocks = 148 * occs[oi];
      printf("blocks/SM %d x 256 thr: ", occs[oi]);
      printf(" ld U4 %.1f", run<0, 4>(idx, tab, out, n, blocks, 256));
      printf(" ld U8 %.1f", run<0, 8>(idx, tab, out, n, blocks, 256));
      printf(" ldg U8 %.1f", run<1, 8>(idx, tab, out, n, blocks, 256));
      printf(" ldcg U8 %.1f", run<2, 8>(idx, tab, out, n, blocks, 256));
      printf(" ldg8B U8 %.1f", run<3, 8>(idx, tab, out, n, blocks, 256));
      printf(" ld U16 %.1f\n", run<0, 16>(idx, tab, out, n, blocks, 256));
   }
   // sorted-ish indices: locality upper bound
   for( long long i = 0; i < n; ++i )
      h[i] = (int)((i / 10) % ncols);
   cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
   printf("sequential indices (10 per column): ld U8 %.1f us\n", run<0, 8>(idx, tab, out, n, 148 * 8, 256));
   return 0;
}
