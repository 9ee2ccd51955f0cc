// Texture-unit filtering probe: dumps hardware tex1D<float4>/tex3D<float> linear-filter
// results for random normalised coordinates so the software texture unit in oracle/
// (and the shared-memory TF lookup in the product kernel) can be fitted to the hardware.
// Build: nvcc -arch=sm_100a -o texprobe texprobe.cu ; run on the GPU box, writes gpurun_out/texprobe_*.bin
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__global__ void k1d(cudaTextureObject_t t, const float* u, float4* out, int n){
  int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n) out[i]=tex1D<float4>(t,u[i]);
}
__global__ void k3d(cudaTextureObject_t t, const float* u, float* out, int n){
  int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n) out[i]=tex3D<float>(t,u[3*i],u[3*i+1],u[3*i+2]);
}
static uint32_t rs=12345u; static float rnd(){ rs=rs*1664525u+1013904223u; return (rs>>8)*(1.0f/16777216.0f);} 
static void dump(const char* name,const void* p,size_t bytes){ FILE* f=fopen(name,"wb"); fwrite(p,1,bytes,f); fclose(f);} 
int main(){
  const int N=1<<20;
  // ---- 1D float4 table, 256 texels
  {
    std::vector<float> tab(256*4); for(auto&v:tab) v=rnd();
    cudaArray_t arr; auto desc=cudaCreateChannelDesc(32,32,32,32,cudaChannelFormatKindFloat);
    CK(cudaMallocArray(&arr,&desc,256));
    CK(cudaMemcpy2DToArray(arr,0,0,tab.data(),256*16,256*16,1,cudaMemcpyHostToDevice));
    cudaResourceDesc rd{}; rd.resType=cudaResourceTypeArray; rd.res.array.array=arr;
    cudaTextureDesc td{}; td.addressMode[0]=cudaAddressModeClamp; td.filterMode=cudaFilterModeLinear; td.readMode=cudaReadModeElementType; td.normalizedCoords=1;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex,&rd,&td,nullptr));
    std::vector<float> u(N); for(int i=0;i<N;i++){ u[i]= (i<4096)? (i/4096.0f) : rnd()*1.02f-0.01f; }
    float *du; float4* dout; CK(cudaMalloc(&du,N*4)); CK(cudaMalloc(&dout,N*16));
    CK(cudaMemcpy(du,u.data(),N*4,cudaMemcpyHostToDevice));
    k1d<<<(N+255)/256,256>>>(tex,du,dout,N); CK(cudaDeviceSynchronize());
    std::vector<float> out(N*4); CK(cudaMemcpy(out.data(),dout,N*16,cudaMemcpyDeviceToHost));
    dump("gpurun_out/texprobe_1d_tab.bin",tab.data(),tab.size()*4);
    dump("gpurun_out/texprobe_1d_u.bin",u.data(),u.size()*4);
    dump("gpurun_out/texprobe_1d_out.bin",out.data(),out.size()*4);
  }
  // ---- 3D float volume 16x12x10 (non-cubic to expose axis order), and a 1024-wide one for large-N precision
  for(int pass=0;pass<2;pass++){
    int nx= pass?1024:16, ny= pass?8:12, nz= pass?4:10;
    std::vector<float> vol((size_t)nx*ny*nz); for(auto&v:vol) v=rnd();
    cudaArray_t arr; auto desc=cudaCreateChannelDesc(32,0,0,0,cudaChannelFormatKindFloat);
    CK(cudaMalloc3DArray(&arr,&desc,make_cudaExtent(nx,ny,nz)));
    cudaMemcpy3DParms cp{}; cp.srcPtr=make_cudaPitchedPtr(vol.data(),nx*4,nx,ny); cp.dstArray=arr; cp.extent=make_cudaExtent(nx,ny,nz); cp.kind=cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&cp));
    cudaResourceDesc rd{}; rd.resType=cudaResourceTypeArray; rd.res.array.array=arr;
    cudaTextureDesc td{}; for(int a=0;a<3;a++) td.addressMode[a]=cudaAddressModeClamp; td.filterMode=cudaFilterModeLinear; td.readMode=cudaReadModeElementType; td.normalizedCoords=1;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex,&rd,&td,nullptr));
    std::vector<float> u((size_t)N*3); for(auto&v:u) v=rnd()*1.04f-0.02f;
    float *du,*dout; CK(cudaMalloc(&du,(size_t)N*12)); CK(cudaMalloc(&dout,N*4));
    CK(cudaMemcpy(du,u.data(),(size_t)N*12,cudaMemcpyHostToDevice));
    k3d<<<(N+255)/256,256>>>(tex,du,dout,N); CK(cudaDeviceSynchronize());
    std::vector<float> out(N); CK(cudaMemcpy(out.data(),dout,N*4,cudaMemcpyDeviceToHost));
    char nm[128];
    snprintf(nm,128,"gpurun_out/texprobe_3d%d_vol.bin",pass); dump(nm,vol.data(),vol.size()*4);
    snprintf(nm,128,"gpurun_out/texprobe_3d%d_u.bin",pass); dump(nm,u.data(),u.size()*4);
    snprintf(nm,128,"gpurun_out/texprobe_3d%d_out.bin",pass); dump(nm,out.data(),out.size()*4);
  }
  printf("texprobe done\n");
  return 0;
}
