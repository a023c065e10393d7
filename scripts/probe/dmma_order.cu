// Which FMA order does DMMA.8x8x4 (mma.sync.m8n8k4.f64) reproduce?  Compares bits with candidate orders on random data.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884( double& d0, double& d1, double a, double b ) {
  asm volatile( "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"( d0 ), "+d"( d1 ) : "d"( a ), "d"( b ) );
}
__global__ void k( const double* A, const double* B, const double* C, double* D, int trials ) {
  const int lane = threadIdx.x, fr = lane >> 2, fk = lane & 3;
  for ( int t = 0; t < trials; t++ ) {
    const double* a = A + t * 32; const double* b = B + t * 32; const double* c = C + t * 64;
    double d0 = c[fr * 8 + 2 * fk], d1 = c[fr * 8 + 2 * fk + 1];
    dmma884( d0, d1, a[fr * 4 + fk], b[fr * 4 + fk] );   // a[m][k], b[n][k]
    D[t * 64 + fr * 8 + 2 * fk] = d0; D[t * 64 + fr * 8 + 2 * fk + 1] = d1;
  }
}
int main() {
  const int T = 4096;
  double *hA = new double[T * 32], *hB = new double[T * 32], *hC = new double[T * 64], *hD = new double[T * 64];
  srand( 1 );
  auto rnd = []() { return ( rand() / ( double )RAND_MAX - 0.5 ) * ( ( rand() % 7 ) ? 1.0 : 1e6 ); };
  for ( int i = 0; i < T * 32; i++ ) { hA[i] = rnd(); hB[i] = rnd(); }
  for ( int i = 0; i < T * 64; i++ ) hC[i] = rnd();
  double *A, *B, *C, *D;
  cudaMalloc( &A, T * 32 * 8 ); cudaMalloc( &B, T * 32 * 8 ); cudaMalloc( &C, T * 64 * 8 ); cudaMalloc( &D, T * 64 * 8 );
  cudaMemcpy( A, hA, T * 32 * 8, cudaMemcpyHostToDevice ); cudaMemcpy( B, hB, T * 32 * 8, cudaMemcpyHostToDevice ); cudaMemcpy( C, hC, T * 64 * 8, cudaMemcpyHostToDevice );
  k<<<1, 32>>>( A, B, C, D, T );
  cudaMemcpy( hD, D, T * 64 * 8, cudaMemcpyDeviceToHost );
  long mism[4] = { 0, 0, 0, 0 };
  for ( int t = 0; t < T; t++ ) for ( int m = 0; m < 8; m++ ) for ( int n = 0; n < 8; n++ ) {
    const double* a = hA + t * 32 + m * 4; const double* b = hB + t * 32 + n * 4; const double c = hC[t * 64 + m * 8 + n], d = hD[t * 64 + m * 8 + n];
    double asc = c; for ( int q = 0; q < 4; q++ ) asc = __builtin_fma( a[q], b[q], asc );
    double desc = c; for ( int q = 3; q >= 0; q-- ) desc = __builtin_fma( a[q], b[q], desc );
    double pair = c + ( __builtin_fma( a[1], b[1], a[0] * b[0] ) + __builtin_fma( a[3], b[3], a[2] * b[2] ) );
    double last = __builtin_fma( a[3], b[3], __builtin_fma( a[2], b[2], __builtin_fma( a[1], b[1], a[0] * b[0] ) ) ) + c;
    if ( memcmp( &asc, &d, 8 ) ) mism[0]++; if ( memcmp( &desc, &d, 8 ) ) mism[1]++; if ( memcmp( &pair, &d, 8 ) ) mism[2]++; if ( memcmp( &last, &d, 8 ) ) mism[3]++;
  }
  printf( "mismatches of %d: fma chain k ascending from c: %ld | descending: %ld | pairwise: %ld | products first then +c: %ld\n", T * 64, mism[0], mism[1], mism[2], mism[3] );
  return 0;
}
