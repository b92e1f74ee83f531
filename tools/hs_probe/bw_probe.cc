#include <thread>
#include <vector>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstdlib>
int main(int argc,char**argv){int T=atoi(argv[1]); size_t n=64<<20; std::vector<std::vector<char>> a(T,std::vector<char>(n,1)); 
 auto t0=std::chrono::steady_clock::now(); std::vector<std::thread> th; std::vector<long> s(T*16);
 for(int t=0;t<T;t++) th.emplace_back([&,t]{long x=0; const long* p=(const long*)a[t].data(); for(int r=0;r<4;r++) for(size_t i=0;i<n/8;i++) x+=p[i]; s[t*16]=x;});
 for(auto&x:th)x.join(); auto t1=std::chrono::steady_clock::now(); double ms=std::chrono::duration<double,std::milli>(t1-t0).count();
 printf("T=%d %.1f GB/s (%ld)\n",T,4.0*T*n/ms/1e6,s[0]);}
