// C entry points over the Options class of votca_b200/host/gwbse.h (device-free part) for tests/test_options_cpu.py.
#include <cstring>
#include <string>

#include "../../votca_b200/host/gwbse.h"

using namespace votca::xtp;

extern "C" {
// RPA::UpdateRPAInputEnergies (rpa.cc:32-70) is pure host arithmetic: run it on an RPA object whose device handle
// is never touched (the Device behind the reference is deliberately not constructed - there is no GPU here).
int rpa_update_energies(long homo, long rpamin, long rpamax, const double* dft, long ndft, const double* gw, long ngw,
                        long qpmin, double* out) {
  try {
    alignas(Device) static unsigned char no_device[sizeof(Device)];
    const Device& dev = *reinterpret_cast<const Device*>(no_device);
    Logger log;
    TCMatrix_gwbse Mmn(dev);
    RPA rpa(log, Mmn);
    rpa.configure(homo, rpamin, rpamax);
    rpa.UpdateRPAInputEnergies(VectorXd(dft, ndft), VectorXd(gw, ngw), qpmin);
    const VectorXd& e = rpa.getRPAInputEnergies();
    for (long i = 0; i < e.size(); ++i) out[i] = e(i);
    return (int)e.size();
  } catch (...) {
    return -1;
  }
}

void* opt_new() { return new Options(); }
void opt_free(void* o) { delete static_cast<Options*>(o); }
int opt_load_xml(void* o, const char* path) {
  try {
    static_cast<Options*>(o)->LoadFromXML(path);
    return 0;
  } catch (...) {
    return 1;
  }
}
int opt_set(void* o, const char* key, const char* value) {
  try {
    static_cast<Options*>(o)->set(key, value);
    return 0;
  } catch (...) {
    return 1;
  }
}
// 0: ok, 1: key unknown / empty
int opt_get(void* o, const char* key, char* out, int cap) {
  try {
    Options* opt = static_cast<Options*>(o);
    if (!opt->exists(key)) return 1;
    const std::string v = opt->str(key);
    std::strncpy(out, v.c_str(), cap - 1);
    out[cap - 1] = 0;
    return 0;
  } catch (...) {
    return 1;
  }
}
}
