// Minimal GGUF v3 reader for the device-resident forward: header, metadata
// key/values and tensor infos of an in-memory image.  Written from the format
// as the reference parses it (gguf.cpp:258-304: header, kv pairs, tensor infos,
// data section aligned to a hard-coded 32; value types gguf.h:15-29); only what
// Model needs (model.cpp:58-238).  Host-side plumbing, no arithmetic.
#pragma once

#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace llmi {

struct GgufValue {
  uint32_t type = 0;
  uint64_t u = 0;     // any integer / bool
  double f = 0.0;     // any float
  std::string s;      // string
  std::vector<GgufValue> arr;
};

struct GgufTensor {
  std::string name;
  std::vector<uint64_t> shape;
  uint32_t type = 0;
  uint64_t offset = 0;
  const uint8_t* data = nullptr;  // into the caller's image
  uint64_t n_elements() const {
    uint64_t n = 1;
    for (uint64_t d : shape) n *= d;
    return n;
  }
};

class GgufImage {
 public:
  GgufImage(const uint8_t* data, uint64_t size) : p_(data), size_(size) { parse(); }

  const GgufValue* find(const std::string& key) const {
    auto it = kv_.find(key);
    return it == kv_.end() ? nullptr : &it->second;
  }
  const GgufTensor* tensor(const std::string& name) const {
    auto it = tensors_.find(name);
    return it == tensors_.end() ? nullptr : &it->second;
  }
  const std::map<std::string, GgufTensor>& tensors() const { return tensors_; }

 private:
  const uint8_t* p_;
  uint64_t size_, pos_ = 0;
  std::map<std::string, GgufValue> kv_;
  std::map<std::string, GgufTensor> tensors_;

  template <typename T>
  T rd() {
    if (pos_ + sizeof(T) > size_) throw std::runtime_error("Read beyond end of file");
    T v;
    memcpy(&v, p_ + pos_, sizeof(T));
    pos_ += sizeof(T);
    return v;
  }
  std::string rd_str() {
    const uint64_t n = rd<uint64_t>();
    if (n > size_ || pos_ + n > size_) throw std::runtime_error("String length exceeds file size");
    std::string s(reinterpret_cast<const char*>(p_ + pos_), n);
    pos_ += n;
    return s;
  }
  GgufValue rd_value(uint32_t t) {
    GgufValue v;
    v.type = t;
    switch (t) {
      case 0: v.u = rd<uint8_t>(); break;
      case 1: v.u = uint64_t(int64_t(rd<int8_t>())); break;
      case 2: v.u = rd<uint16_t>(); break;
      case 3: v.u = uint64_t(int64_t(rd<int16_t>())); break;
      case 4: v.u = rd<uint32_t>(); break;
      case 5: v.u = uint64_t(int64_t(rd<int32_t>())); break;
      case 6: v.f = rd<float>(); break;
      case 7: v.u = rd<uint8_t>() ? 1 : 0; break;
      case 8: v.s = rd_str(); break;
      case 9: {
        const uint32_t et = rd<uint32_t>();
        const uint64_t n = rd<uint64_t>();
        v.u = n;
        // big string arrays (the 262k-entry vocabulary) are skipped, not stored
        const bool keep = et != 8 || n <= 64;
        for (uint64_t i = 0; i < n; ++i) {
          GgufValue e = rd_value(et);
          if (keep) v.arr.push_back(std::move(e));
        }
      } break;
      case 10: v.u = rd<uint64_t>(); break;
      case 11: v.u = uint64_t(rd<int64_t>()); break;
      case 12: v.f = rd<double>(); break;
      default: throw std::runtime_error("Unsupported GGUF value type");
    }
    return v;
  }
  void parse() {
    const uint32_t magic = rd<uint32_t>();
    if (magic != 0x46554747u) throw std::runtime_error("Invalid GGUF magic number");
    (void)rd<uint32_t>();  // version
    const uint64_t n_tensors = rd<uint64_t>();
    const uint64_t n_kv = rd<uint64_t>();
    for (uint64_t i = 0; i < n_kv; ++i) {
      std::string k = rd_str();
      const uint32_t t = rd<uint32_t>();
      kv_[k] = rd_value(t);
    }
    std::vector<GgufTensor> infos;
    for (uint64_t i = 0; i < n_tensors; ++i) {
      GgufTensor t;
      t.name = rd_str();
      const uint32_t nd = rd<uint32_t>();
      for (uint32_t d = 0; d < nd; ++d) t.shape.push_back(rd<uint64_t>());
      t.type = rd<uint32_t>();
      t.offset = rd<uint64_t>();
      infos.push_back(std::move(t));
    }
    const uint64_t start = (pos_ + 31) & ~uint64_t(31);
    for (auto& t : infos) {
      if (start + t.offset > size_) throw std::runtime_error("Read beyond end of file");
      t.data = p_ + start + t.offset;
      tensors_[t.name] = t;
    }
  }
};

}  // namespace llmi
