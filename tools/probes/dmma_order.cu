// Does mma.sync.m8n8k4.f64 round like an ascending-k FMA chain  fma(a3,b3,fma(a2,b2,fma(a1,b1,fma(a0,b0,c)))) ?
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const double* A, const double* B, double* D, int ntiles)
{
  int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const double* a = A + tile * 32; const double* b = B + tile * 32;
    double a0 = a[g * 4 + t];          // A row g, col t
    double b0 = b[t * 8 + g];          // B row t (k), col g (n)
    double d0 = 0.0, d1 = 0.0;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a0), "d"(b0));
    D[tile * 64 + g * 8 + 2 * t] = d0; D[tile * 64 + g * 8 + 2 * t + 1] = d1;
  }
}
int main()
{
  const int nt = 4096; size_t na = nt * 32, nd = nt * 64;
  double *hA = (double*)malloc(na * 8), *hB = (double*)malloc(na * 8), *hD = (double*)malloc(nd * 8);
  srand(7);
  for (size_t i = 0; i < na; ++i) { hA[i] = ldexp((double)rand() / RAND_MAX, -(rand() % 40)); hB[i] = (double)rand() / RAND_MAX; }
  double *dA, *dB, *dD; cudaMalloc(&dA, na * 8); cudaMalloc(&dB, na * 8); cudaMalloc(&dD, nd * 8);
  cudaMemcpy(dA, hA, na * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, na * 8, cudaMemcpyHostToDevice);
  k<<<64, 32>>>(dA, dB, dD, nt); cudaMemcpy(hD, dD, nd * 8, cudaMemcpyDeviceToHost);
  long asc = 0, desc = 0, mulfirst = 0, pair = 0, tot = 0;
  for (int tile = 0; tile < nt; ++tile) for (int m = 0; m < 8; ++m) for (int n = 0; n < 8; ++n) {
    const double* a = hA + tile * 32 + m * 4; const double* b = hB + tile * 32;
    double p[4]; for (int kk = 0; kk < 4; ++kk) p[kk] = b[kk * 8 + n];
    double r_asc = fma(a[3], p[3], fma(a[2], p[2], fma(a[1], p[1], fma(a[0], p[0], 0.0))));
    double r_desc = fma(a[0], p[0], fma(a[1], p[1], fma(a[2], p[2], fma(a[3], p[3], 0.0))));
    double r_pair = (fma(a[1], p[1], a[0] * p[0])) + (fma(a[3], p[3], a[2] * p[2]));
    double d = hD[tile * 64 + m * 8 + n];
    ++tot; asc += (d == r_asc); desc += (d == r_desc); pair += (d == r_pair);
  }
  printf("tiles %d: matches ascending-FMA-chain %ld / %ld, descending %ld, pairwise %ld\n", nt, asc, tot, desc, pair);
  return 0;
}
