#include <cstdio>
#include <cuda_runtime.h>
template<int ILP> __global__ void k(double* out, int iters, double m, double b){
  double a[ILP];
  for(int i=0;i<ILP;i++) a[i]=threadIdx.x*1e-3+i;
  for(int it=0; it<iters; ++it){
#pragma unroll
    for(int i=0;i<ILP;i++) a[i]=fma(a[i],m,b);
  }
  double s=0; for(int i=0;i<ILP;i++) s+=a[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP> void run(int warps_per_sm){
  int sms=148; int threads=32*warps_per_sm; // one block per SM
  double* out; cudaMalloc(&out, sizeof(double)*sms*threads);
  int iters=1<<16;
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<ILP><<<sms,threads>>>(out,iters,0.999999,1e-9); cudaDeviceSynchronize();
  cudaEventRecord(a); k<ILP><<<sms,threads>>>(out,iters,0.999999,1e-9); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  double fmas=(double)iters*ILP*sms*threads;
  printf("warps/SM=%2d ILP=%d  %.2f TFLOP/s  (%.1f cycles per dependent step per warp)\n", warps_per_sm, ILP, 2*fmas/(ms*1e-3)/1e12, ms*1e-3*1.965e9/iters);
  cudaFree(out);
}
int main(){
  for(int w: {4,8,16,32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
