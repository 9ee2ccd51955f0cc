// fuzz_importers.cpp — mutation fuzzer for the volume-file importers (include/dvr_import.h): reads a seed file, applies
// 1..FUZZ_MAXMUT random mutations (byte / bit flips, cuts, insertions, truncation, extreme 64-bit integers), writes the
// case next to the binary and imports it with dvr_import_volume; an imported payload is touched like a consumer would.
// Build with the sanitizers and link the importer's source directly:
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -Iinclude -o /tmp/fuzz/fuzz \
//       tools/fuzz_importers.cpp visrtx_b200/importers/volume_import.cpp -lz
//   ASAN_OPTIONS=allocator_may_return_null=1 /tmp/fuzz/fuzz seed.nvdb .nvdb 2500 <rng seed>
// Round 2: ~45 k cases over 24 VTI encodings, 5 NanoVDB files (none / zip / raw buffer / quantised / two grids) and MHD:
// no ASan / UBSan report; one finding — a segment's declared grid size was allocated before it was held against the
// file (hundreds of GB requested) — fixed in importNvdb and pinned by
// tests/test_importers_host.py::test_import_nvdb_holds_the_declared_grid_size_against_what_the_file_can_deliver.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "dvr_import.h"
static std::vector<unsigned char> readAll(const char *p){ std::vector<unsigned char> v; FILE*f=fopen(p,"rb"); if(!f) return v; fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); v.resize(n); if(n) fread(v.data(),1,n,f); fclose(f); return v; }
int main(int argc,char**argv){
  if(argc<4) return 1;
  const char*seed=argv[1]; const char*ext=argv[2]; int iters=atoi(argv[3]); unsigned s=argc>4?atoi(argv[4]):1;
  auto base=readAll(seed); if(base.empty()) return 2;
  std::mt19937 rng(s);
  std::string out=std::string(argv[0])+"_case_"+std::to_string(s)+ext;
  int ok=0,bad=0;
  for(int it=0;it<iters;++it){
    auto v=base;
    int maxm = getenv("FUZZ_MAXMUT") ? atoi(getenv("FUZZ_MAXMUT")) : 8; int nm=1+rng()%maxm;
    for(int m=0;m<nm;++m){
      int kind=rng()%6; size_t pos=v.empty()?0:rng()%v.size();
      if(kind==0&&!v.empty()) v[pos]=(unsigned char)rng();
      else if(kind==1&&!v.empty()) v[pos]^=(unsigned char)(1u<<(rng()%8));
      else if(kind==2&&v.size()>8){ size_t n=1+rng()%std::min<size_t>(v.size()-pos,64); v.erase(v.begin()+pos,v.begin()+pos+n);} 
      else if(kind==3){ size_t n=1+rng()%32; std::vector<unsigned char> ins(n); for(auto&c:ins)c=(unsigned char)rng(); v.insert(v.begin()+pos,ins.begin(),ins.end()); }
      else if(kind==4&&v.size()>16){ v.resize(rng()%v.size()); }
      else if(kind==5&&v.size()>8){ // overwrite 8 bytes with an extreme integer
        unsigned long long x[4]={0xffffffffffffffffull,0x8000000000000000ull,0x7fffffffull,0x100000000ull}; unsigned long long val=x[rng()%4]; size_t p2=std::min(pos,v.size()-8); memcpy(&v[p2],&val,8);} 
    }
    FILE*f=fopen(out.c_str(),"wb"); if(!f) return 3; if(!v.empty()) fwrite(v.data(),1,v.size(),f); fclose(f);
    DvrVolumeFile vf; memset(&vf,0,sizeof(vf));
    int rc=dvr_import_volume(out.c_str(),&vf);
    if(rc==0){ ok++; // touch the payload like a consumer would
      volatile unsigned char acc=0; const unsigned char*d=(const unsigned char*)vf.data; if(d&&vf.bytes){ acc^=d[0]; acc^=d[vf.bytes-1]; acc^=d[vf.bytes/2]; }
      dvr_import_free(&vf);} else bad++;
  }
  printf("%s: %d iterations, %d imported, %d rejected\n",ext,iters,ok,bad); return 0; }
