// tools/plan_stats.cpp - prints the size of the static schedule of a matrix: value slots, tasks,
// levels, multiply-adds, 128-lane steps and chunks of the multi-warp form, peak live slots, and
// (4th argument) the number of tasks per level.
//
//   python tools/dump_sax_pattern.py 128
//   cd klujax_b200/csrc && g++ -O2 -std=c++17 -I. ../../tools/plan_stats.cpp \
//       ordering.cpp seed_lu.cpp plan.cpp -o /tmp/plan_stats
//   /tmp/plan_stats /tmp/pl/sax128.txt 111 1 x      # split mode, program mode (1 = factor + solve)
#include <cstdio>
#include <cstdlib>
#include <complex>
#include <vector>
#include "symbolic.hpp"
#include "plan.hpp"
using namespace kb;
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"r"); int n,nz; fscanf(f,"%d %d",&n,&nz);
  std::vector<int> Ai(nz),Aj(nz); std::vector<std::complex<double>> Ax(nz);
  for(int e=0;e<nz;++e){double re,im; fscanf(f,"%d %d %lf %lf",&Ai[e],&Aj[e],&re,&im); Ax[e]={re,im};}
  int split=argc>2?atoi(argv[2]):111; int mode=argc>3?atoi(argv[3]):PM_FACTOR_SOLVE;
  CscPattern A=coo_to_csc(n,nz,Ai.data(),Aj.data());
  Ordering ord; analyze_pattern(A,ord);
  std::vector<std::complex<double>> cv(nz); for(int p=0;p<nz;++p) cv[p]=Ax[A.coo_of[p]];
  PivotPlan piv; if(!seed_factor<std::complex<double>>(A,ord,cv.data(),piv)){printf("seed fail\n");return 1;}
  SlotLayout lay; build_layout(A,ord,piv,lay);
  TaskProgram tp; build_program(ord,piv,lay,mode,1,tp,split, 8);
  PackedProgram pp; bool ok=pack_program_wide(lay,tp,pp);
  printf("n_val %d tasks %zu levels %d mac %lld steps %d chunks %d ok %d\n",lay.n_val,tp.tasks.size(),tp.nlevels,(long long)tp.n_mac,pp.nbatch,pp.nchunks,ok);
  {
    int nv=lay.n_val; std::vector<int> birth(nv,1<<30), death(nv,-1);
    for(int s2=0;s2<nv;++s2) if(lay.slot_csc[s2]>=0) birth[s2]=0;
    for(auto&t:tp.tasks){ auto touch=[&](int s2,int l){ if(s2<nv){ birth[s2]=std::min(birth[s2],l); death[s2]=std::max(death[s2],l);} };
      touch(t.out,t.level); if(t.div>=0) touch(t.div,t.level); for(int q=0;q<t.plen;++q){touch(tp.pa[t.pbeg+q],t.level);touch(tp.pb[t.pbeg+q],t.level);} }
    std::vector<int> delta(tp.nlevels+3,0); int never=0;
    for(int s2=0;s2<nv;++s2){ if(death[s2]<0){never++; continue;} delta[birth[s2]]++; delta[death[s2]+1]--; }
    int live=0,peak=0,pl=0; for(int l=0;l<=tp.nlevels+1;++l){ live+=delta[l]; if(live>peak){peak=live;pl=l;} }
    printf("slots %d never-touched %d peak live %d at level %d\n",nv,never,peak,pl);
  }
  // per level: tasks
  if(argc>4){ for(int l=0;l<tp.nlevels;++l) printf("%d ",tp.level_ptr[l+1]-tp.level_ptr[l]); printf("\n"); }
}
