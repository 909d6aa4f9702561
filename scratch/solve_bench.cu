#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <cublas_v2.h>
#define CK(x) do{ auto e=(x); if(e!=0){printf("err %d line %d\n",(int)e,__LINE__); exit(1);} }while(0)
int main(){
  const int n=496, B=64;
  std::vector<double> h((size_t)n*n);
  // SPD: A = G G^T/n + I
  std::vector<double> G((size_t)n*n); srand(1); for(auto&v:G) v=rand()/(double)RAND_MAX-0.5;
  for(int i=0;i<n;i++) for(int j=0;j<n;j++){ double s=0; for(int k=0;k<n;k++) s+=G[(size_t)i*n+k]*G[(size_t)j*n+k]; h[(size_t)i*n+j]=s/n+(i==j?1.0:0.0);}
  double *dA0,*dA,*db; int *info; CK(cudaMalloc(&dA0,sizeof(double)*n*n*B)); CK(cudaMalloc(&dA,sizeof(double)*n*n*B)); CK(cudaMalloc(&db,sizeof(double)*n*B)); CK(cudaMalloc(&info,sizeof(int)*B));
  for(int b=0;b<B;b++) CK(cudaMemcpy(dA0+(size_t)b*n*n,h.data(),sizeof(double)*n*n,cudaMemcpyHostToDevice));
  cusolverDnHandle_t cs; CK(cusolverDnCreate(&cs)); cublasHandle_t cb; CK(cublasCreate(&cb));
  int lwork; CK(cusolverDnDpotrf_bufferSize(cs,CUBLAS_FILL_MODE_LOWER,n,dA,n,&lwork)); double* work; CK(cudaMalloc(&work,sizeof(double)*lwork));
  std::vector<double*> hp(B), hb(B); for(int b=0;b<B;b++){hp[b]=dA+(size_t)b*n*n; hb[b]=db+(size_t)b*n;}
  double **dp,**dbp; CK(cudaMalloc(&dp,sizeof(double*)*B)); CK(cudaMalloc(&dbp,sizeof(double*)*B)); CK(cudaMemcpy(dp,hp.data(),sizeof(double*)*B,cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbp,hb.data(),sizeof(double*)*B,cudaMemcpyHostToDevice));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  auto reset=[&](){ CK(cudaMemcpy(dA,dA0,sizeof(double)*n*n*B,cudaMemcpyDeviceToDevice)); CK(cudaMemset(db,0,sizeof(double)*n*B)); };
  for(int rep=0;rep<3;rep++){
    reset(); cudaEventRecord(e0); for(int i=0;i<10;i++){ CK(cusolverDnDpotrf(cs,CUBLAS_FILL_MODE_LOWER,n,dA+(size_t)i*n*n,n,work,lwork,info)); } cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); printf("potrf single: %.1f us each\n",ms*100);
    cudaEventRecord(e0); for(int i=0;i<10;i++){ CK(cusolverDnDpotrs(cs,CUBLAS_FILL_MODE_LOWER,n,1,dA+(size_t)i*n*n,n,db,n,info)); } cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); printf("potrs single: %.1f us each\n",ms*100);
    reset(); cudaEventRecord(e0); CK(cusolverDnDpotrfBatched(cs,CUBLAS_FILL_MODE_LOWER,n,dp,n,info,1)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); printf("potrfBatched(1): %.1f us\n",ms*1000);
    reset(); cudaEventRecord(e0); CK(cusolverDnDpotrfBatched(cs,CUBLAS_FILL_MODE_LOWER,n,dp,n,info,B)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); printf("potrfBatched(64): %.1f us total\n",ms*1000);
    cudaEventRecord(e0); CK(cusolverDnDpotrsBatched(cs,CUBLAS_FILL_MODE_LOWER,n,1,dp,n,dbp,n,info,1)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); printf("potrsBatched(1): %.1f us\n",ms*1000);
    const double one=1.0;
    cudaEventRecord(e0); for(int i=0;i<10;i++){ CK(cublasDtrsv(cb,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_N,CUBLAS_DIAG_NON_UNIT,n,dA,n,db,1)); CK(cublasDtrsv(cb,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_T,CUBLAS_DIAG_NON_UNIT,n,dA,n,db,1)); } cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); printf("2x cublasDtrsv: %.1f us\n",ms*100);
    cudaEventRecord(e0); for(int i=0;i<10;i++){ CK(cublasDtrsm(cb,CUBLAS_SIDE_LEFT,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_N,CUBLAS_DIAG_NON_UNIT,n,1,&one,dA,n,db,n)); CK(cublasDtrsm(cb,CUBLAS_SIDE_LEFT,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_T,CUBLAS_DIAG_NON_UNIT,n,1,&one,dA,n,db,n)); } cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); printf("2x cublasDtrsm: %.1f us\n",ms*100);
  }
  // big n
  return 0;
}
