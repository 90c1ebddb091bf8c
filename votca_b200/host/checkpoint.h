// CheckpointFile / CheckpointWriter - writes the GW-BSE results in the .orb (HDF5) layout of the reference
// (SURVEY.md 8f, N2), without an HDF5 library (none exists in this environment).
//
// Mirrors the call surface of xtp/include/votca/xtp/checkpointwriter.h:49-353 and checkpoint.h:
//   CheckpointFile cpf(filename);                       // CheckpointAccessLevel::CREATE
//   CheckpointWriter w = cpf.getWriter("/QMdata");
//   w(value, "name");  w.openChild("BSE_singlet");
// with the same mapping of C++ types to HDF5 objects:
//   fundamental types, bool (as Index), std::string  -> attribute with a 1-element simple dataspace
//                                                       (WriteScalar, checkpointwriter.h:165-200; strings are
//                                                       variable-length, stored in a global heap collection)
//   matrices / vectors                                 -> 2-D IEEE_F64LE dataset (rows, max(cols,1)) in C order
//                                                       (WriteData, checkpointwriter.h:203-253)
//   std::vector<Vector3d>                              -> group with datasets ind0, ind1, ... (:297-311)
//   tools::EigenSystem                                 -> group with eigenvalues / eigenvectors / eigenvectors2
//                                                       datasets and an "info" attribute (:313-328)
// File format: HDF5 superblock version 2, version-2 object headers, compact link and attribute storage, contiguous
// datasets - the message encodings are the ones libhdf5 itself wrote into the reference's checked-in .orb files
// (xtp/src/tests/DataFiles/xtp_tools_integration_tests/*.orb); all checksums are Jenkins lookup3, as the format
// specifies.  Compound tables (atoms, basis shells: CptTable) are DFT-side inputs and are not produced here.
#pragma once
#include <cstdint>
#include <cstring>
#include <ctime>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "matrix.h"

namespace votca {
namespace xtp {

namespace cpt_detail {

inline uint32_t rot(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

// Bob Jenkins' lookup3 hashlittle(), byte-wise form (public domain); the HDF5 metadata checksum
inline uint32_t lookup3(const uint8_t* k, size_t length, uint32_t initval = 0) {
  uint32_t a, b, c;
  a = b = c = 0xdeadbeefu + (uint32_t)length + initval;
  auto w = [](const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); };
  while (length > 12) {
    a += w(k);
    b += w(k + 4);
    c += w(k + 8);
    a -= c; a ^= rot(c, 4); c += b;
    b -= a; b ^= rot(a, 6); a += c;
    c -= b; c ^= rot(b, 8); b += a;
    a -= c; a ^= rot(c, 16); c += b;
    b -= a; b ^= rot(a, 19); a += c;
    c -= b; c ^= rot(b, 4); b += a;
    length -= 12;
    k += 12;
  }
  if (length == 0) return c;
  uint8_t t[12] = {0};
  std::memcpy(t, k, length);
  a += w(t);
  b += w(t + 4);
  c += w(t + 8);
  c ^= b; c -= rot(b, 14);
  a ^= c; a -= rot(c, 11);
  b ^= a; b -= rot(a, 25);
  c ^= b; c -= rot(b, 16);
  a ^= c; a -= rot(c, 4);
  b ^= a; b -= rot(a, 14);
  c ^= b; c -= rot(b, 24);
  return c;
}

struct Bytes {
  std::vector<uint8_t> d;
  void u8(uint8_t v) { d.push_back(v); }
  void u16(uint16_t v) { for (int i = 0; i < 2; ++i) d.push_back(uint8_t(v >> (8 * i))); }
  void u32(uint32_t v) { for (int i = 0; i < 4; ++i) d.push_back(uint8_t(v >> (8 * i))); }
  void u64(uint64_t v) { for (int i = 0; i < 8; ++i) d.push_back(uint8_t(v >> (8 * i))); }
  void raw(const void* p, size_t n) { const uint8_t* q = static_cast<const uint8_t*>(p); d.insert(d.end(), q, q + n); }
  void raw(const Bytes& o) { d.insert(d.end(), o.d.begin(), o.d.end()); }
  void hex(const char* s) {
    auto nib = [](char c) { return uint8_t(c <= '9' ? c - '0' : (c | 32) - 'a' + 10); };
    for (; s[0] && s[1]; s += 2) d.push_back(uint8_t(nib(s[0]) << 4 | nib(s[1])));
  }
  size_t size() const { return d.size(); }
};

constexpr uint64_t UNDEF = 0xFFFFFFFFFFFFFFFFull;

// datatype messages as libhdf5 encodes them (read from the reference's .orb files)
inline void dtype_f64(Bytes& b) { b.hex("11203f000800000000004000340b0034ff030000"); }  // IEEE_F64LE
inline void dtype_i32(Bytes& b) { b.hex("100800000400000000002000"); }                  // NATIVE_INT
inline void dtype_i64(Bytes& b) { b.hex("100800000800000000004000"); }                  // NATIVE_LONG
inline void dtype_u8(Bytes& b) { b.hex("100000000100000000000800"); }                   // NATIVE_UINT8
inline void dtype_vlstr(Bytes& b) { b.hex("1901000010000000100000000100000000000800"); }  // variable-length string

struct Attr {
  std::string name;
  enum Kind { I32, I64, U8, F64, STR } kind;
  int64_t i = 0;
  double f = 0.0;
  std::string s;
  uint32_t heap_index = 0;
};

struct Node {
  std::string name;
  bool is_group = true;
  std::vector<std::unique_ptr<Node>> children;
  std::vector<Attr> attrs;
  uint64_t rows = 0, cols = 0;
  std::vector<double> data;  // C order

  Node* child(const std::string& n) {
    for (auto& c : children)
      if (c->name == n) return c.get();
    return nullptr;
  }
};

}  // namespace cpt_detail

class CheckpointFile;

class CheckpointWriter {
 public:
  // fundamental types (bool is written as Index, checkpointwriter.h:97-110)
  void operator()(int v, const std::string& name) const { num(name, cpt_detail::Attr::I32, v, 0.0); }
  void operator()(long v, const std::string& name) const { num(name, cpt_detail::Attr::I64, v, 0.0); }
  void operator()(long long v, const std::string& name) const { num(name, cpt_detail::Attr::I64, v, 0.0); }
  void operator()(std::uint8_t v, const std::string& name) const { num(name, cpt_detail::Attr::U8, v, 0.0); }
  void operator()(bool v, const std::string& name) const { num(name, cpt_detail::Attr::I64, v ? 1 : 0, 0.0); }
  void operator()(double v, const std::string& name) const { num(name, cpt_detail::Attr::F64, 0, v); }
  void operator()(const std::string& v, const std::string& name) const {
    cpt_detail::Attr& a = attr(name);
    a.kind = cpt_detail::Attr::STR;
    a.s = v;
  }
  void operator()(const char* v, const std::string& name) const { (*this)(std::string(v), name); }

  // Eigen matrices / vectors (vectors are n x 1 matrices; zero columns are stored as one)
  void operator()(const MatrixXd& m, const std::string& name) const {
    cpt_detail::Node& d = dataset(name);
    d.rows = (uint64_t)m.rows();
    d.cols = m.cols() == 0 ? 1 : (uint64_t)m.cols();  // "eigen vectors are n,1 matrices", checkpointwriter.h:211-215
    d.data.assign((size_t)(d.rows * d.cols), 0.0);
    for (Index i = 0; i < m.rows(); ++i)
      for (Index j = 0; j < m.cols(); ++j) d.data[(size_t)i * m.cols() + j] = m(i, j);
  }
  void operator()(const VectorXd& v, const std::string& name) const {
    cpt_detail::Node& d = dataset(name);
    d.rows = (uint64_t)v.size();
    d.cols = 1;
    d.data.assign(v.data(), v.data() + v.size());
  }
  // std::vector<Eigen::Vector3d>
  void operator()(const std::vector<VectorXd>& v, const std::string& name) const {
    CheckpointWriter parent = openChild(name);
    for (size_t c = 0; c < v.size(); ++c) parent(v[c], "ind" + std::to_string(c));
  }
  // tools::EigenSystem
  void WriteEigenSystem(const VectorXd& eigenvalues, const MatrixXd& eigenvectors, const MatrixXd& eigenvectors2,
                        long info, const std::string& name) const {
    CheckpointWriter parent = openChild(name);
    parent(eigenvalues, "eigenvalues");
    parent(eigenvectors, "eigenvectors");
    parent(eigenvectors2, "eigenvectors2");
    parent(info, "info");
  }

  CheckpointWriter openChild(const std::string& childName) const {
    cpt_detail::Node* c = node_->child(childName);
    if (c && !c->is_group) throw std::runtime_error("Could not open or create " + path_ + "/" + childName);
    if (!c) {
      node_->children.emplace_back(new cpt_detail::Node);
      c = node_->children.back().get();
      c->name = childName;
    }
    return CheckpointWriter(c, path_ + "/" + childName);
  }

 private:
  friend class CheckpointFile;
  CheckpointWriter(cpt_detail::Node* n, std::string path) : node_(n), path_(std::move(path)) {}
  cpt_detail::Attr& attr(const std::string& name) const {
    if (name.empty() || name.size() > 200) throw std::runtime_error("Could not write " + name + " to " + path_);
    for (auto& a : node_->attrs)
      if (a.name == name) return a;  // reopened and overwritten, as createAttribute/openAttribute does
    node_->attrs.emplace_back();
    node_->attrs.back().name = name;
    return node_->attrs.back();
  }
  void num(const std::string& name, cpt_detail::Attr::Kind k, int64_t i, double f) const {
    cpt_detail::Attr& a = attr(name);
    a.kind = k;
    a.i = i;
    a.f = f;
  }
  cpt_detail::Node& dataset(const std::string& name) const {
    if (name.empty() || name.size() > 200) throw std::runtime_error("Could not write " + name + " to " + path_);
    cpt_detail::Node* c = node_->child(name);
    if (c && c->is_group) throw std::runtime_error("Could not write " + name + " to " + path_);
    if (!c) {
      node_->children.emplace_back(new cpt_detail::Node);
      c = node_->children.back().get();
      c->name = name;
      c->is_group = false;
    }
    return *c;
  }
  cpt_detail::Node* node_;
  std::string path_;
};

class CheckpointFile {
 public:
  explicit CheckpointFile(const std::string& filename) : filename_(filename) {}
  ~CheckpointFile() {
    if (!closed_) {
      try {
        Close();
      } catch (...) {
      }
    }
  }
  std::string getFileName() const { return filename_; }

  CheckpointWriter getWriter(const std::string& path = "/") {
    CheckpointWriter w(&root_, "");
    size_t p = 0;
    while (p < path.size()) {
      const size_t q = path.find('/', p);
      const std::string part = path.substr(p, q == std::string::npos ? std::string::npos : q - p);
      if (!part.empty()) w = w.openChild(part);
      if (q == std::string::npos) break;
      p = q + 1;
    }
    return w;
  }

  // serialises the tree; the file is complete only after Close()
  void Close() {
    using namespace cpt_detail;
    closed_ = true;
    Bytes out;
    out.d.assign(48, 0);  // superblock, filled in last
    // ---- global heap collection with every string attribute ------------------------------------------------
    std::vector<Attr*> strings;
    collect_strings(root_, strings);
    uint64_t gcol_addr = 0;
    if (!strings.empty()) {
      gcol_addr = out.size();
      Bytes objs;
      uint32_t idx = 0;
      for (Attr* a : strings) {
        a->heap_index = ++idx;
        if (idx > 0xFFFF) throw std::runtime_error("too many string attributes in " + filename_);
        objs.u16((uint16_t)idx);
        objs.u16(0);
        objs.u32(0);
        objs.u64(a->s.size());
        objs.raw(a->s.data(), a->s.size());
        while (objs.size() % 8) objs.u8(0);
      }
      uint64_t total = 16 + objs.size() + 16;  // header + objects + free-space object header
      total = total < 4096 ? 4096 : (total + 4095) / 4096 * 4096;
      out.raw("GCOL", 4);
      out.u8(1);
      out.u8(0), out.u8(0), out.u8(0);
      out.u64(total);
      out.raw(objs);
      const uint64_t rest = total - 16 - objs.size();  // object 0: the free space, its own header included
      out.u16(0), out.u16(0), out.u32(0);
      out.u64(rest);
      out.d.resize(out.size() + (size_t)(rest - 16), 0);
    }
    const uint64_t root_addr = write_node(root_, out, gcol_addr);
    // ---- superblock version 2 ------------------------------------------------------------------------------
    Bytes sb;
    sb.hex("894844460d0a1a0a");
    sb.u8(2), sb.u8(8), sb.u8(8), sb.u8(0);
    sb.u64(0);      // base address
    sb.u64(UNDEF);  // superblock extension
    sb.u64(out.size());
    sb.u64(root_addr);
    sb.u32(lookup3(sb.d.data(), sb.size()));
    std::memcpy(out.d.data(), sb.d.data(), 48);
    std::ofstream fh(filename_, std::ios::binary | std::ios::trunc);
    if (!fh) throw std::runtime_error("Could not write " + filename_);
    fh.write(reinterpret_cast<const char*>(out.d.data()), (std::streamsize)out.size());
    if (!fh) throw std::runtime_error("Could not write " + filename_);
  }

 private:
  static void collect_strings(cpt_detail::Node& n, std::vector<cpt_detail::Attr*>& out) {
    for (auto& a : n.attrs)
      if (a.kind == cpt_detail::Attr::STR) out.push_back(&a);
    for (auto& c : n.children) collect_strings(*c, out);
  }

  static void message(cpt_detail::Bytes& chunk, uint8_t type, uint8_t flags, const cpt_detail::Bytes& body) {
    chunk.u8(type);
    chunk.u16((uint16_t)body.size());
    chunk.u8(flags);
    chunk.raw(body);
  }

  static void attribute_messages(const cpt_detail::Node& n, cpt_detail::Bytes& chunk, uint64_t gcol_addr) {
    using namespace cpt_detail;
    if (n.attrs.empty()) return;
    Bytes ainfo;  // attribute info: no dense storage
    ainfo.u8(0), ainfo.u8(0), ainfo.u64(UNDEF), ainfo.u64(UNDEF);
    message(chunk, 0x15, 0x04, ainfo);
    for (const Attr& a : n.attrs) {
      Bytes dt, data;
      switch (a.kind) {
        case Attr::I32: dtype_i32(dt); data.u32((uint32_t)(int32_t)a.i); break;
        case Attr::I64: dtype_i64(dt); data.u64((uint64_t)a.i); break;
        case Attr::U8: dtype_u8(dt); data.u8((uint8_t)a.i); break;
        case Attr::F64: dtype_f64(dt); data.raw(&a.f, 8); break;
        case Attr::STR:
          dtype_vlstr(dt);
          data.u32((uint32_t)a.s.size());
          data.u64(gcol_addr);
          data.u32(a.heap_index);
          break;
      }
      Bytes sp;  // simple dataspace, rank 1, dims {1}, max dims {1}
      sp.u8(2), sp.u8(1), sp.u8(1), sp.u8(1), sp.u64(1), sp.u64(1);
      Bytes body;
      body.u8(3), body.u8(0);
      body.u16((uint16_t)(a.name.size() + 1));
      body.u16((uint16_t)dt.size());
      body.u16((uint16_t)sp.size());
      body.u8(0);  // ASCII
      body.raw(a.name.c_str(), a.name.size() + 1);
      body.raw(dt);
      body.raw(sp);
      body.raw(data);
      message(chunk, 0x0C, 0, body);
    }
  }

  // writes children first (their addresses go into the link messages), returns the object header address
  static uint64_t write_node(const cpt_detail::Node& n, cpt_detail::Bytes& out, uint64_t gcol_addr) {
    using namespace cpt_detail;
    Bytes chunk;
    if (n.is_group) {
      std::vector<uint64_t> addr;
      for (const auto& c : n.children) addr.push_back(write_node(*c, out, gcol_addr));
      Bytes linfo;  // link info: no dense storage
      linfo.u8(0), linfo.u8(0), linfo.u64(UNDEF), linfo.u64(UNDEF);
      message(chunk, 0x02, 0, linfo);
      Bytes ginfo;
      ginfo.u8(0), ginfo.u8(0);
      message(chunk, 0x0A, 0x01, ginfo);
      for (size_t i = 0; i < n.children.size(); ++i) {
        const std::string& nm = n.children[i]->name;
        Bytes link;
        link.u8(1), link.u8(0);  // version 1, 1-byte name length, hard link
        link.u8((uint8_t)nm.size());
        link.raw(nm.data(), nm.size());
        link.u64(addr[i]);
        message(chunk, 0x06, 0, link);
      }
    } else {
      uint64_t data_addr = UNDEF;
      const uint64_t nbytes = 8ull * n.data.size();
      if (nbytes) {
        while (out.size() % 8) out.u8(0);
        data_addr = out.size();
        out.raw(n.data.data(), (size_t)nbytes);
      }
      Bytes sp;  // simple dataspace, rank 2, with max dims
      sp.u8(2), sp.u8(2), sp.u8(1), sp.u8(1);
      sp.u64(n.rows), sp.u64(n.cols), sp.u64(n.rows), sp.u64(n.cols);
      message(chunk, 0x01, 0, sp);
      Bytes dt;
      dtype_f64(dt);
      message(chunk, 0x03, 0x01, dt);
      Bytes fill;
      fill.u8(3), fill.u8(0x0a);
      message(chunk, 0x05, 0x01, fill);
      Bytes layout;  // version 3, contiguous
      layout.u8(3), layout.u8(1), layout.u64(data_addr), layout.u64(nbytes);
      message(chunk, 0x08, 0, layout);
    }
    attribute_messages(n, chunk, gcol_addr);
    // ---- version-2 object header: times stored, 4-byte chunk size ------------------------------------------
    while (out.size() % 8) out.u8(0);
    const uint64_t at = out.size();
    Bytes oh;
    oh.raw("OHDR", 4);
    oh.u8(2);
    oh.u8(0x22);
    const uint32_t now = (uint32_t)std::time(nullptr);
    for (int i = 0; i < 4; ++i) oh.u32(now);
    oh.u32((uint32_t)chunk.size());
    oh.raw(chunk);
    oh.u32(lookup3(oh.d.data(), oh.size()));
    out.raw(oh);
    return at;
  }

  std::string filename_;
  cpt_detail::Node root_;
  bool closed_ = false;
};

}  // namespace xtp
}  // namespace votca
