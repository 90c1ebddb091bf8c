// FP64 peak probes for B200: DMMA shapes, DFMA, cuBLAS Dgemm.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__device__ __forceinline__ void mma884(double (&c)[2], double a, double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n":"+d"(c[0]),"+d"(c[1]):"d"(a),"d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], double a0,double a1, double b){
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n":"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a0),"d"(a1),"d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double* a, const double* b){
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n":"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(b[0]),"d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double* a, const double* b){
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n":"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),"d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
}

template<int SHAPE, int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double seed){
  double a[8], b[4];
  for(int i=0;i<8;i++) a[i]=seed+threadIdx.x*1e-9+i;
  for(int i=0;i<4;i++) b[i]=seed*0.5+i;
  double c4[NACC][4]; double c2[NACC][2];
  for(int j=0;j<NACC;j++){ for(int i=0;i<4;i++) c4[j][i]=0; c2[j][0]=c2[j][1]=0; }
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int j=0;j<NACC;j++){
      if(SHAPE==0) mma884(c2[j], a[0], b[0]);
      if(SHAPE==1) mma1684(c4[j], a[0], a[1], b[0]);
      if(SHAPE==2) mma1688(c4[j], a, b);
      if(SHAPE==3) mma16816(c4[j], a, b);
    }
  }
  double s=0; for(int j=0;j<NACC;j++){ for(int i=0;i<4;i++) s+=c4[j][i]; s+=c2[j][0]+c2[j][1]; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double seed){
  double a=seed+threadIdx.x*1e-9, b=seed*0.5;
  double c[NACC]; for(int j=0;j<NACC;j++) c[j]=j;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int j=0;j<NACC;j++) c[j]=fma(a,b,c[j]);
  }
  double s=0; for(int j=0;j<NACC;j++) s+=c[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<typename F> float timeit(F f){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<5;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms; }
  return best;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("dev %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int nsm=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double)*nsm*8*256*4));
  const int iters=20000;
  for(int bps=1;bps<=4;bps*=2){
    int grid=nsm*bps;
    #define RUN(SH,NACC,FLOPS,name) { float ms=timeit([&]{ k_dmma<SH,NACC><<<grid,256>>>(out,iters,1.0); }); double fl=(double)grid*8*(double)iters*NACC*FLOPS; printf("%-12s nacc=%d blocks/SM=%d warps/SM=%d : %.2f TFLOP/s (%.3f ms)\n", name,NACC,bps,bps*8, fl/ms/1e9, ms); }
    RUN(0,8,2*8*8*4,"m8n8k4");
    RUN(1,4,2*16*8*4,"m16n8k4");
    RUN(1,8,2*16*8*4,"m16n8k4");
    RUN(2,4,2*16*8*8,"m16n8k8");
    RUN(2,8,2*16*8*8,"m16n8k8");
    RUN(3,4,2*16*8*16,"m16n8k16");
    RUN(3,8,2*16*8*16,"m16n8k16");
    { float ms=timeit([&]{ k_dfma<16><<<grid,256>>>(out,iters,1.0); }); double fl=(double)grid*256*(double)iters*16*2; printf("%-12s blocks/SM=%d : %.2f TFLOP/s (%.3f ms)\n","dfma",bps, fl/ms/1e9, ms); }
  }
  // cuBLAS dgemm
  cublasHandle_t h; cublasCreate(&h);
  for(int n : {4096, 8192}){
    double *A,*B,*C; size_t bytes=sizeof(double)*n*n; CK(cudaMalloc(&A,bytes)); CK(cudaMalloc(&B,bytes)); CK(cudaMalloc(&C,bytes));
    CK(cudaMemset(A,0,bytes)); CK(cudaMemset(B,0,bytes));
    // fill with something non-trivial
    double al=1.0, be=0.0;
    for (auto tr : {CUBLAS_OP_N, CUBLAS_OP_T}) {
      float ms=timeit([&]{ cublasDgemm(h,tr,CUBLAS_OP_N,n,n,n,&al,A,n,B,n,&be,C,n); });
      printf("cublasDgemm %s n=%d: %.2f TFLOP/s (%.3f ms)\n", tr==CUBLAS_OP_N?"NN":"TN", n, 2.0*n*n*n/ms/1e9, ms);
    }
    // sustained 3 s
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int reps = n==8192? 60: 200;
    cudaEventRecord(e0); for(int r=0;r<reps;r++) cublasDgemm(h,CUBLAS_OP_T,CUBLAS_OP_N,n,n,n,&al,A,n,B,n,&be,C,n); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1); printf("cublasDgemm TN n=%d sustained %d reps: %.2f TFLOP/s (%.1f ms total)\n", n, reps, 2.0*n*n*n*reps/ms/1e9, ms);
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  return 0;
}
