// CPU build of the DFT codelets (test infrastructure): lets pytest compare them with numpy.fft.
#include "../../ad-yolo_b200/csrc/fft_codelets.cuh"
using namespace ady;
extern "C" {
void emu_dft16(float* io) { cx<float> x[16]; for (int i=0;i<16;++i) x[i]={io[2*i],io[2*i+1]}; dft16(x); for (int i=0;i<16;++i){io[2*i]=x[i].re;io[2*i+1]=x[i].im;} }
void emu_dft48(float* io) { cx<float> x[48]; for (int i=0;i<48;++i) x[i]={io[2*i],io[2*i+1]}; dft48(x); for (int i=0;i<48;++i){io[2*i]=x[i].re;io[2*i+1]=x[i].im;} }
void emu_dft25(float* io) { cx<float> x[25]; for (int i=0;i<25;++i) x[i]={io[2*i],io[2*i+1]}; dft25(x); for (int i=0;i<25;++i){io[2*i]=x[i].re;io[2*i+1]=x[i].im;} }
void emu_dft48_f64(double* io) { cx<double> x[48]; for (int i=0;i<48;++i) x[i]={io[2*i],io[2*i+1]}; dft48(x); for (int i=0;i<48;++i){io[2*i]=x[i].re;io[2*i+1]=x[i].im;} }
void emu_dft25_f64(double* io) { cx<double> x[25]; for (int i=0;i<25;++i) x[i]={io[2*i],io[2*i+1]}; dft25(x); for (int i=0;i<25;++i){io[2*i]=x[i].re;io[2*i+1]=x[i].im;} }
// full 1200-point PFA on the CPU (48 x 25), natural order in/out
void emu_fft1200(const float* in, float* out) {
  static cx<float> mid[48][25];
  for (int n2=0;n2<25;++n2){ cx<float> x[48]; for(int n1=0;n1<48;++n1){int n=pfa_in(n1,n2); x[n1]={in[2*n],in[2*n+1]};} dft48(x); for(int k1=0;k1<48;++k1) mid[k1][n2]=x[k1]; }
  for (int k1=0;k1<48;++k1){ cx<float> y[25]; for(int n2=0;n2<25;++n2) y[n2]=mid[k1][n2]; dft25(y); for(int k2=0;k2<25;++k2){int k=pfa_out(k1,k2); out[2*k]=y[k2].re; out[2*k+1]=y[k2].im;} }
}
}
