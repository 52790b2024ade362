// tools/bank_conflict_model.cpp - offline model of the shared-memory bank conflicts of the
// multi-warp kernel's schedule (kernels_wide.cuh).  Walks the packed stream the way a warp does
// (groups of 8 lanes per 16-byte wavefront; destination, first and second operand are separate
// accesses) and counts wavefronts.  It reproduced ncu: 28.7 % replays predicted, 28 % measured.
//
//   python tools/dump_sax_pattern.py 128            # writes /tmp/pl/sax128.txt (n, nz, COO triplets)
//   cd klujax_b200/csrc && g++ -O2 -std=c++17 -I. ../../tools/bank_conflict_model.cpp \
//       ordering.cpp seed_lu.cpp plan.cpp -o /tmp/bank_model
//   /tmp/bank_model /tmp/pl/sax128.txt 8            # 2nd argument: compaction bank columns (0 = no compaction)
#include <cstdio>
#include <cstdlib>
#include <complex>
#include <vector>
#include <set>
#include <map>
#include "symbolic.hpp"
#include "plan.hpp"
using namespace kb;
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"r"); int n,nz; if(fscanf(f,"%d %d",&n,&nz)!=2) return 1;
  std::vector<int> Ai(nz),Aj(nz); std::vector<std::complex<double>> Ax(nz);
  for(int e=0;e<nz;++e){double re,im; if(fscanf(f,"%d %d %lf %lf",&Ai[e],&Aj[e],&re,&im)!=4) return 1; Ax[e]={re,im};}
  CscPattern A=coo_to_csc(n,nz,Ai.data(),Aj.data());
  Ordering ord; analyze_pattern(A,ord);
  std::vector<std::complex<double>> cv(nz); for(int p=0;p<nz;++p) cv[p]=Ax[A.coo_of[p]];
  PivotPlan piv; seed_factor<std::complex<double>>(A,ord,cv.data(),piv);
  SlotLayout lay; build_layout(A,ord,piv,lay);
  TaskProgram tp; build_program(ord,piv,lay,PM_FACTOR_SOLVE,1,tp,111, argc>2?atoi(argv[2]):0); printf("n_val %d\n", tp.n_val);
  PackedProgram pp; pack_program_wide(lay,tp,pp,16);
  // walk the stream
  long ideal=0, actual=0; int G=8; long cid[3]={0,0,0}, cac[3]={0,0,0};
  for(int c=0;c<pp.nchunks;++c){ const uint32_t* p=pp.stream.data()+(size_t)c*kChunkWords;
    for(int s=pp.chunk_ptr[c]; s<pp.chunk_ptr[c+1]; ++s){ int P=pp.ctrl[s]&0xF;
      for(int g=0; g<128/G; ++g){
        for(int cls=0; cls<3; ++cls){ 
          std::map<int,std::set<int>> col; int act=0;
          for(int l=g*G;l<(g+1)*G;++l){ uint32_t h=p[2*l]; if(!(h&0x7)) continue; int slot;
            if(cls==0) slot=h>>18; else { uint32_t w=p[2*l+1]; slot= cls==1? (w>>18) : ((w>>4)&0x3FFF); }
            col[slot%G].insert(slot); act++; }
          if(!act) continue; size_t mx=0; for(auto&kv:col) mx=std::max(mx,kv.second.size());
          ideal+=1; actual+=mx; cid[cls]+=1; cac[cls]+=mx; if(cls==0){ideal+=1; actual+=mx;} /* out: load + store */ }
      }
      p+=256; } }
  for(int c=0;c<3;++c) printf("class %d ideal %ld actual %ld\n",c,cid[c],cac[c]); printf("group wavefronts: ideal %ld actual %ld (%.1f%% replays)\n",ideal,actual,100.0*(actual-ideal)/actual);
}
