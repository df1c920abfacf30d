// Minimal safetensors reader (host-only) for cz_model_load_safetensors.
// Replaces VarBuilder::from_mmaped_safetensors (src/models.rs:55) / candle_core::safetensors::load (src/models.rs:137).
// Format: u64 LE header length, JSON header {"name": {"dtype": "BF16", "shape": [..], "data_offsets": [a, b]}, ...},
// raw little-endian tensor data.  Supports F32 / BF16 / F16.  Tensors the model does not know are skipped.
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/candlezip_b200.h"

namespace cz {
void set_error(const std::string &msg);
}

namespace {

struct Cursor {
  const char *p, *end;
  void ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++;
  }
  bool eat(char c) {
    ws();
    if (p < end && *p == c) {
      p++;
      return true;
    }
    return false;
  }
  bool str(std::string &out) {
    ws();
    if (p >= end || *p != '"') return false;
    p++;
    out.clear();
    while (p < end && *p != '"') {
      if (*p == '\\' && p + 1 < end) {
        p++;
        switch (*p) {
          case 'n': out.push_back('\n'); break;
          case 't': out.push_back('\t'); break;
          case 'u': p += 4; out.push_back('?'); break;  // names are ASCII in practice
          default: out.push_back(*p);
        }
        p++;
      } else {
        out.push_back(*p++);
      }
    }
    if (p >= end) return false;
    p++;
    return true;
  }
  bool num(uint64_t &v) {
    ws();
    if (p >= end || *p < '0' || *p > '9') return false;
    v = 0;
    while (p < end && *p >= '0' && *p <= '9') v = v * 10 + (uint64_t)(*p++ - '0');
    return true;
  }
  // skip any JSON value
  bool skip() {
    ws();
    if (p >= end) return false;
    if (*p == '"') {
      std::string s;
      return str(s);
    }
    if (*p == '{' || *p == '[') {
      const char open = *p, close = open == '{' ? '}' : ']';
      p++;
      ws();
      if (eat(close)) return true;
      for (;;) {
        if (open == '{') {
          std::string k;
          if (!str(k) || !eat(':')) return false;
        }
        if (!skip()) return false;
        if (eat(',')) continue;
        return eat(close);
      }
    }
    while (p < end && *p != ',' && *p != '}' && *p != ']') p++;
    return true;
  }
};

struct Entry {
  std::string name, dtype;
  std::vector<uint64_t> shape;
  uint64_t begin = 0, end = 0;
};

bool parse_header(const char *js, size_t n, std::vector<Entry> &out) {
  Cursor c{js, js + n};
  if (!c.eat('{')) return false;
  if (c.eat('}')) return true;
  for (;;) {
    Entry e;
    if (!c.str(e.name) || !c.eat(':')) return false;
    if (e.name == "__metadata__") {
      if (!c.skip()) return false;
    } else {
      if (!c.eat('{')) return false;
      for (;;) {
        std::string k;
        if (!c.str(k) || !c.eat(':')) return false;
        if (k == "dtype") {
          if (!c.str(e.dtype)) return false;
        } else if (k == "shape" || k == "data_offsets") {
          if (!c.eat('[')) return false;
          std::vector<uint64_t> v;
          if (!c.eat(']')) {
            for (;;) {
              uint64_t x;
              if (!c.num(x)) return false;
              v.push_back(x);
              if (c.eat(',')) continue;
              if (!c.eat(']')) return false;
              break;
            }
          }
          if (k == "shape") e.shape = v;
          else if (v.size() == 2) {
            e.begin = v[0];
            e.end = v[1];
          } else return false;
        } else if (!c.skip()) return false;
        if (c.eat(',')) continue;
        if (!c.eat('}')) return false;
        break;
      }
      out.push_back(std::move(e));
    }
    if (c.eat(',')) continue;
    return c.eat('}');
  }
}

}  // namespace

extern "C" int cz_model_load_safetensors(cz_model *m, const char *const *paths, int n_paths) {
  if (!m || !paths || n_paths <= 0) return CZ_ERR_INVALID;
  int loaded = 0;
  for (int f = 0; f < n_paths; f++) {
    int fd = open(paths[f], O_RDONLY);
    if (fd < 0) {
      cz::set_error(std::string("cannot open ") + paths[f]);
      return CZ_ERR_IO;
    }
    struct stat sb;
    fstat(fd, &sb);
    const size_t len = (size_t)sb.st_size;
    void *map = len >= 8 ? mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0) : MAP_FAILED;
    close(fd);
    if (map == MAP_FAILED) {
      cz::set_error(std::string("cannot map ") + paths[f]);
      return CZ_ERR_IO;
    }
    const uint8_t *b = (const uint8_t *)map;
    uint64_t hlen = 0;
    memcpy(&hlen, b, 8);
    std::vector<Entry> entries;
    if (hlen > len - 8 || !parse_header((const char *)b + 8, (size_t)hlen, entries)) {
      munmap(map, len);
      cz::set_error(std::string("bad safetensors header in ") + paths[f]);
      return CZ_ERR_FORMAT;
    }
    const uint8_t *data = b + 8 + hlen;
    const size_t data_len = len - 8 - (size_t)hlen;
    const int n_slots = cz_model_tensor_count(m);
    for (const Entry &e : entries) {
      int dt = e.dtype == "F32" ? CZ_DTYPE_F32 : e.dtype == "BF16" ? CZ_DTYPE_BF16 : e.dtype == "F16" ? CZ_DTYPE_F16 : -1;
      size_t n = 1;
      for (uint64_t d : e.shape) n *= (size_t)d;
      bool known = false;
      for (int i = 0; i < n_slots && !known; i++) {
        const char *nm;
        size_t ne;
        cz_model_tensor_info(m, i, &nm, &ne);
        known = e.name == nm;
      }
      if (!known) continue;  // e.g. rotary inv_freq buffers
      const size_t esz = dt == CZ_DTYPE_F32 ? 4 : 2;
      if (dt < 0 || e.end < e.begin || e.end > data_len || e.end - e.begin != n * esz) {
        munmap(map, len);
        cz::set_error("unsupported dtype or bad offsets for tensor " + e.name);
        return CZ_ERR_FORMAT;
      }
      int rc = cz_model_set_tensor(m, e.name.c_str(), data + e.begin, dt, n);
      if (rc != CZ_OK) {
        munmap(map, len);
        return rc;
      }
      loaded++;
    }
    munmap(map, len);
  }
  if (loaded == 0) {
    cz::set_error("no known tensors found in the safetensors files");
    return CZ_ERR_FORMAT;
  }
  return CZ_OK;
}
