#include "../../flowgnn_b200/csrc/host_stage.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
using namespace fg;
int main(int argc,char**argv){
  int T = argc>1?atoi(argv[1]):8;
  size_t N=1034135,E=2253418;
  std::vector<int32_t> f(9*N), e(2*E), a(3*E);
  for(size_t i=0;i<f.size();i++) f[i]=i%119; for(size_t i=0;i<e.size();i++) e[i]=i%60; for(size_t i=0;i<a.size();i++) a[i]=i%2;
  NarrowPlan p; p.layout(N,E,true);
  std::vector<uint8_t> blk(p.bytes);
  HostPool pool(T); NarrowJob job;
  for(int r=0;r<8;r++){
    auto t0=std::chrono::steady_clock::now();
    job.start(pool,p,f.data(),e.data(),a.data(),blk.data()); bool ok[3]; job.finish(pool,ok);
    auto t1=std::chrono::steady_clock::now();
    printf("T=%d %.3f ms ok=%d%d%d bytes %zu\n",T,std::chrono::duration<double,std::milli>(t1-t0).count(),ok[0],ok[1],ok[2],p.bytes);
  }
  // check
  for(size_t i=0;i<f.size();i++) if(blk[p.off_feat+i]!=(uint8_t)f[i]){puts("BAD f");return 1;}
  for(size_t i=0;i<e.size();i++) if(((uint16_t*)(blk.data()+p.off_edge))[i]!=(uint16_t)e[i]){puts("BAD e");return 1;}
  for(size_t i=0;i<a.size();i++) if(blk[p.off_attr+i]!=(uint8_t)a[i]){puts("BAD a");return 1;}
  puts("ok");
}
