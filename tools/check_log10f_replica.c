#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static const double T[16][2]={
  { 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 },
  { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
  { 0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2 },
  { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
  { 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 },
  { 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3 },
  { 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 },
  { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
  { 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 },
  { 0x1p+0, 0x0p+0 },
  { 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5 },
  { 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4 },
  { 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3 },
  { 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3 },
  { 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2 },
  { 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 }};
static const double LN2=0x1.62e42fefa39efp-1, A0=-0x1.00ea348b88334p-2, A1=0x1.5575b0be00b6ap-2, A2=-0x1.ffffef20a4123p-2;
static inline uint32_t asu(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float asf(uint32_t u){float f;memcpy(&f,&u,4);return f;}
#ifdef USEFMA
#define MAD(a,b,c) fma(a,b,c)
#else
#define MAD(a,b,c) ((a)*(b)+(c))
#endif
static float my_logf(float x){
  uint32_t ix=asu(x); if(ix==0x3f800000) return 0;
  uint32_t tmp=ix-0x3f330000; int i=(tmp>>19)%16; int k=(int32_t)tmp>>23; uint32_t iz=ix-(tmp&0x1ffu<<23);
  double invc=T[i][0],logc=T[i][1],z=asf(iz);
  double r=MAD(z,invc,-1.0); double y0=MAD((double)k,LN2,logc);
  double r2=r*r; double y=MAD(A1,r,A2); y=MAD(A0,r2,y); y=MAD(y,r2,(y0+r)); return (float)y; }
static float my_log10f(float x){
  const float ivln10=4.3429449201e-01f, log10_2hi=3.0102920532e-01f, log10_2lo=7.9034151668e-07f;
  int32_t hx=asu(x),k=0,i; k+=(hx>>23)-127; i=((uint32_t)k&0x80000000)>>31; hx=(hx&0x007fffff)|((0x7f-i)<<23);
  float y=(float)(k+i); x=asf(hx);
  float z=y*log10_2lo+ivln10*my_logf(x); return z+y*log10_2hi; }
int main(){
  for(int i=0;i<16;i++){ double lc=-log(T[i][0]); printf("%d %a %a %s\n",i,lc,T[i][1], lc==T[i][1]?"ok":"DIFF"); }
  long bad=0,tot=0,badl=0; for(uint32_t u=asu(1e-6f);u<asu(1e9f);u++){ float x=asf(u); if(asu(log10f(x))!=asu(my_log10f(x))) bad++; if(asu(logf(x))!=asu(my_logf(x))) badl++; tot++; }
  printf("tot %ld bad log10f %ld bad logf %ld\n",tot,bad,badl); return 0; }
