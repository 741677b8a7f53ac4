#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <omp.h>
static const double C0=0x1p0, C1=-0x1.ffffffd0c621cp-2, C2=0x1.55553e1068f19p-5, C3=-0x1.6c087e89a359dp-10, C4=0x1.99343027bf8c3p-16;
static const double S1=-0x1.555545995a603p-3, S2=0x1.1107605230bc4p-7, S3=-0x1.994eb3774cf24p-13;
static const double HPI_INV=0x1.45F306DC9C883p+23, HPI=0x1.921FB54442D18p0, PI63=0x1.921FB54442D18p-62;
static const uint32_t INVPIO4[24]={0xa2,0xa2f9,0xa2f983,0xa2f9836e,0xf9836e4e,0x836e4e44,0x6e4e4415,0x4e441529,0x441529fc,0x1529fc27,0x29fc2757,0xfc2757d1,0x2757d1f5,0x57d1f534,0xd1f534dd,0xf534ddc0,0x34ddc0db,0xddc0db62,0xc0db6295,0xdb629599,0x6295993c,0x95993c43,0x993c4390,0x3c439041};
static inline uint32_t asu(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline uint32_t top12(float f){return (asu(f)>>20)&0x7ff;}
/* sgn: polynomial sign flip (table[1] = negated coefficients) */
static inline float poly(double x,double x2,int n,int neg){
  double sg = neg?-1.0:1.0;
  if((n&1)==0){ double x3=x*x2; double s1=fma(x2,S3,S2); double x7=x3*x2; double s=fma(x3,S1,x); return fma(x7,s1,s); }
  else { double x4=x2*x2; double c2=fma(x2,sg*C4,sg*C3); double c1=fma(x2,sg*C1,sg*C0); double x6=x4*x2; double c=fma(x4,sg*C2,c1); return fma(x6,c2,c); }
}
static inline double red_fast(double x,int*np){ double r=x*HPI_INV; int n=((int32_t)r+0x800000)>>24; *np=n; return fma(-(double)n,HPI,x); }
static inline double red_large(uint32_t xi,int*np){
  const uint32_t*arr=&INVPIO4[(xi>>26)&15]; int shift=(xi>>23)&7; uint64_t n,res0,res1,res2;
  xi=(xi&0xffffff)|0x800000; xi<<=shift;
  res0=xi*arr[0]; res1=(uint64_t)xi*arr[4]; res2=(uint64_t)xi*arr[8];
  res0=(res2>>32)|(res0<<32); res0+=res1;
  n=(res0+(1ULL<<61))>>62; res0-=n<<62; double x=(int64_t)res0; *np=n; return x*PI63; }
static const double SIGN[4]={1.0,-1.0,-1.0,1.0};
float my_sinf(float y){
  double x=y; int n;
  if(top12(y)<top12(0x1.921fb6p-1f)){ double s=x*x; if(top12(y)<top12(0x1p-12f)) return y; return poly(x,s,0,0);} 
  else if(top12(y)<top12(120.0f)){ x=red_fast(x,&n); double s=SIGN[n&3]; return poly(x*s,x*x,n,(n&2)!=0);} 
  else if(top12(y)<top12(INFINITY)){ uint32_t xi=asu(y); int sign=xi>>31; x=red_large(xi,&n); double s=SIGN[(n+sign)&3]; return poly(x*s,x*x,n,((n+sign)&2)!=0);} 
  return y-y;
}
float my_cosf(float y){
  double x=y; int n;
  if(top12(y)<top12(0x1.921fb6p-1f)){ double s=x*x; if(top12(y)<top12(0x1p-12f)) return 1.0f; return poly(x,s,1,0);} 
  else if(top12(y)<top12(120.0f)){ x=red_fast(x,&n); double s=SIGN[n&3]; return poly(x*s,x*x,n^1,(n&2)!=0);} 
  else if(top12(y)<top12(INFINITY)){ uint32_t xi=asu(y); int sign=xi>>31; x=red_large(xi,&n); double s=SIGN[(n+sign)&3]; return poly(x*s,x*x,n^1,((n+sign)&2)!=0);} 
  return y-y;
}
int main(){
  /* all floats with |x| < 2^17 */
  uint32_t hi=asu(131072.0f); long bad_s=0,bad_c=0,tot=0;
  #pragma omp parallel for reduction(+:bad_s,bad_c,tot) schedule(static)
  for(uint32_t u=0;u<hi;u++){ float f; memcpy(&f,&u,4);
    for(int sg=0;sg<2;sg++){ float x=sg?-f:f; float a=sinf(x),b=my_sinf(x); if(asu(a)!=asu(b)) bad_s++; a=cosf(x); b=my_cosf(x); if(asu(a)!=asu(b)) bad_c++; tot++; } }
  printf("tot %ld bad_sin %ld bad_cos %ld\n",tot,bad_s,bad_c);
  return 0; }
int main_rng(){
  float edges[]={0x1p-12f,0.1f,0.5f,0.785f,0x1.921fb6p-1f,1.0f,1.5f,1.6f,2.0f,3.0f,4.0f,10.0f,119.9f,121.0f,1000.f,1e5f};
  for(int e=0;e+1<sizeof(edges)/4;e++){ long bs=0,bc=0,tot=0; uint32_t lo=asu(edges[e]),hi=asu(edges[e+1]);
   for(uint32_t u=lo;u<hi;u+= (hi-lo>2000000? (hi-lo)/2000000:1)){ float x; memcpy(&x,&u,4); if(asu(sinf(x))!=asu(my_sinf(x))) bs++; if(asu(cosf(x))!=asu(my_cosf(x))) bc++; tot++;}
   printf("[%g,%g) tot %ld bad_sin %ld bad_cos %ld\n",edges[e],edges[e+1],tot,bs,bc);} 
  float x=0.3f; printf("%a %a | %a %a\n", sinf(x), my_sinf(x), cosf(x), my_cosf(x));
  x=1.2f; printf("%a %a | %a %a\n", sinf(x), my_sinf(x), cosf(x), my_cosf(x));
  return 0;}
