// Dumps the exact trilinear tap weights of the texture unit: a 2x2x2 one-hot f32 texture per tap,
// sampled at every (kx,ky,kz)/256 fractional offset with kx,ky,kz in a strided set.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__global__ void k(cudaTextureObject_t t, float* out, int n, int stride, int off){
  int i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=n*n*n) return;
  int kx=(i%n)*stride+off, ky=((i/n)%n)*stride+off, kz=(i/(n*n))*stride+off;
  float ux=(0.5f+kx/256.0f)*0.5f, uy=(0.5f+ky/256.0f)*0.5f, uz=(0.5f+kz/256.0f)*0.5f;
  out[i]=tex3D<float>(t,ux,uy,uz);
}
int main(){
  const int n=64, stride=4; 
  for(int off=0; off<2; off++){
  std::vector<float> all; 
  for(int tap=0;tap<8;tap++){
    float vol[8]={0}; vol[tap]=1.f;  // index = x + 2y + 4z
    cudaArray_t arr; auto desc=cudaCreateChannelDesc(32,0,0,0,cudaChannelFormatKindFloat);
    CK(cudaMalloc3DArray(&arr,&desc,make_cudaExtent(2,2,2)));
    cudaMemcpy3DParms cp{}; cp.srcPtr=make_cudaPitchedPtr(vol,8,2,2); cp.dstArray=arr; cp.extent=make_cudaExtent(2,2,2); cp.kind=cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&cp));
    cudaResourceDesc rd{}; rd.resType=cudaResourceTypeArray; rd.res.array.array=arr;
    cudaTextureDesc td{}; for(int a=0;a<3;a++) td.addressMode[a]=cudaAddressModeClamp; td.filterMode=cudaFilterModeLinear; td.readMode=cudaReadModeElementType; td.normalizedCoords=1;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex,&rd,&td,nullptr));
    float* d; CK(cudaMalloc(&d,n*n*n*4));
    k<<<(n*n*n+255)/256,256>>>(tex,d,n,stride,off*3); CK(cudaDeviceSynchronize());
    std::vector<float> h(n*n*n); CK(cudaMemcpy(h.data(),d,n*n*n*4,cudaMemcpyDeviceToHost));
    all.insert(all.end(),h.begin(),h.end());
    cudaFree(d); cudaDestroyTextureObject(tex); cudaFreeArray(arr);
  }
  char nm[64]; snprintf(nm,64,"gpurun_out/texweights_off%d.bin",off*3);
  FILE* f=fopen(nm,"wb"); fwrite(all.data(),4,all.size(),f); fclose(f);
  }
  printf("texweights done\n"); return 0;
}
