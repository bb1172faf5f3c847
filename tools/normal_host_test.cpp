// Host check of the table-driven Box-Muller in swalbe.jl_b200/csrc/normal.cuh (no GPU needed):
//   g++ -O2 -ffp-contract=off -I swalbe.jl_b200/csrc tools/normal_host_test.cpp -o /tmp/normal_host_test && /tmp/normal_host_test [N]
// Prints the largest error of -2 ln u, cos φ, sin φ against long double, and the moments / tail fractions of 2N deviates.
#include "normal.cuh"
using namespace swalbe;
#include <stdio.h>
#include <stdlib.h>
#include <random>
static void philox(uint32_t c0,uint32_t c1,uint32_t c2,uint32_t c3,uint32_t k0,uint32_t k1,uint32_t out[4]){
  for(int r=0;r<10;++r){ unsigned long long p0=(unsigned long long)0xD2511F53u*c0,p1=(unsigned long long)0xCD9E8D57u*c2;
    uint32_t hi0=p0>>32,lo0=(uint32_t)p0,hi1=p1>>32,lo1=(uint32_t)p1; uint32_t n0=hi1^c1^k0,n1=lo1,n2=hi0^c3^k1,n3=lo0;
    c0=n0;c1=n1;c2=n2;c3=n3;k0+=0x9E3779B9u;k1+=0xBB67AE85u;} out[0]=c0;out[1]=c1;out[2]=c2;out[3]=c3;}
int main(int argc,char**argv){
  static NormalTables T; for(int k=0;k<NRM_LOG_N+NRM_ANG_N;++k) normal_table_entry(T,k);
  const long N=argc>1?atol(argv[1]):20000000; double maxe_ln=0,maxe_c=0,maxe_s=0,maxrel_ln=0; long double s1=0,s2=0,s3=0,s4=0,s6=0,sxy=0; long tail3=0,tail4=0,tail5=0;
  for(long i=0;i<N;++i){ uint32_t r[4]; philox((uint32_t)i,0,11,0,99,5,r);
    if(i<2000){ r[0]= i<1000? (i%33==32?0:(1u<<(i%32))) : r[0]; }   // force deep tails
    double a,c,s; normal_polar_from_bits(r,T,a,c,s);
    // exact reference
    int e=r[0]?__builtin_clz(r[0]):32; if(e==32){uint32_t y=((r[2]&0xfff)<<20)|0x80000u; e+=__builtin_clz(y);}
    long double m=1.0L+(long double)(((unsigned long long)r[1]<<20)|(r[2]>>12))/4503599627370496.0L;
    long double lnu=logl(m)-(e+1)*0.693147180559945309417232121458L;
    long double phi=2.0L*3.14159265358979323846264338327950288L*((long double)r[3]+0.5L)/4294967296.0L;
    double e1=fabs((double)(a-(-2.0L*lnu))); if(e1>maxe_ln)maxe_ln=e1; double rel=e1/fabs((double)(2*lnu)); if(rel>maxrel_ln)maxrel_ln=rel;
    double e2=fabs((double)(c-cosl(phi))),e3=fabs((double)(s-sinl(phi))); if(e2>maxe_c)maxe_c=e2; if(e3>maxe_s)maxe_s=e3;
    if(i>=2000){ double rad=sqrt(a); double z0=rad*c,z1=rad*s; s1+=z0+z1; s2+=z0*z0+z1*z1; s3+=z0*z0*z0+z1*z1*z1; s4+=z0*z0*z0*z0+z1*z1*z1*z1; s6+=pow(z0,6)+pow(z1,6); sxy+=z0*z1;
      tail3+=(fabs(z0)>3)+(fabs(z1)>3); tail4+=(fabs(z0)>4)+(fabs(z1)>4); tail5+=(fabs(z0)>5)+(fabs(z1)>5);} }
  long M=2*(N-2000);
  printf("max abs err -2lnu %.3g (rel %.3g)  cos %.3g  sin %.3g\n",maxe_ln,maxrel_ln,maxe_c,maxe_s);
  printf("mean %.3Lg var %.6Lf skew %.3Lg kurt %.5Lf m6 %.4Lf corr %.3Lg\n",s1/M,s2/M,s3/M,s4/M,s6/M,sxy/(M/2));
  printf("tails: >3 %.5g (exp 2.6998e-3)  >4 %.4g (6.334e-5)  >5 %.3g (5.733e-7)\n",(double)tail3/M,(double)tail4/M,(double)tail5/M);
}
