// C entry points over votca_b200/host/checkpoint.h for the CPU tests (tests/test_checkpoint_writer_cpu.py).
#include <memory>
#include <string>

#include "../../votca_b200/host/checkpoint.h"

using namespace votca;
using namespace votca::xtp;

extern "C" {

void* cpt_open(const char* path) { return new CheckpointFile(path); }

int cpt_close(void* h) {
  std::unique_ptr<CheckpointFile> f(static_cast<CheckpointFile*>(h));
  try {
    f->Close();
    return 0;
  } catch (...) {
    return 1;
  }
}

// kind: 0 int, 1 long, 2 uint8, 3 double, 4 string, 5 bool
int cpt_attr(void* h, const char* group, const char* name, int kind, long i, double d, const char* s) {
  try {
    CheckpointWriter w = static_cast<CheckpointFile*>(h)->getWriter(group);
    switch (kind) {
      case 0: w((int)i, name); break;
      case 1: w((long)i, name); break;
      case 2: w((std::uint8_t)i, name); break;
      case 3: w(d, name); break;
      case 4: w(std::string(s), name); break;
      case 5: w(i != 0, name); break;
      default: return 2;
    }
    return 0;
  } catch (...) {
    return 1;
  }
}

// data column-major rows x cols (Eigen layout); vector != 0 writes a VectorXd
int cpt_dataset(void* h, const char* group, const char* name, long rows, long cols, const double* data, int vector) {
  try {
    CheckpointWriter w = static_cast<CheckpointFile*>(h)->getWriter(group);
    if (vector) {
      w(VectorXd(data, rows), name);
    } else {
      w(MatrixXd(data, rows, cols, rows > 0 ? rows : 1), name);
    }
    return 0;
  } catch (...) {
    return 1;
  }
}

int cpt_eigensystem(void* h, const char* group, const char* name, long n, long k, const double* ev, const double* v1,
                    long k2, const double* v2, long info) {
  try {
    CheckpointWriter w = static_cast<CheckpointFile*>(h)->getWriter(group);
    w.WriteEigenSystem(VectorXd(ev, k), MatrixXd(v1, n, k, n > 0 ? n : 1), MatrixXd(v2, k2 ? n : 0, k2, n > 0 ? n : 1),
                       info, name);
    return 0;
  } catch (...) {
    return 1;
  }
}

int cpt_vec3list(void* h, const char* group, const char* name, long count, const double* xyz) {
  try {
    CheckpointWriter w = static_cast<CheckpointFile*>(h)->getWriter(group);
    std::vector<VectorXd> v;
    for (long c = 0; c < count; ++c) v.emplace_back(xyz + 3 * c, 3);
    w(v, name);
    return 0;
  } catch (...) {
    return 1;
  }
}

int cpt_group(void* h, const char* group) {
  try {
    static_cast<CheckpointFile*>(h)->getWriter(group);
    return 0;
  } catch (...) {
    return 1;
  }
}

unsigned cpt_lookup3(const unsigned char* p, long n) { return cpt_detail::lookup3(p, (size_t)n); }
}
