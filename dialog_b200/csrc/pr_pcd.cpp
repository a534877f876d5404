// pr_pcd.cpp — the cloud reader on the boundary: what pcl::io::loadPCDFile<PointT> gives the reference's callers
// (Dialog/PCLViewer.cpp:80-89, `pcl::io::loadPCDFile(filename, *cloud)` into a PointCloud<PointXYZ>) for PCD v0.7 files:
// DATA ascii, binary and binary_compressed (LZF, field-major), x / y / z anywhere in the record and of any numeric
// type, other fields (rgb, normals, ...) skipped.  The points come back as pcl::PointXYZ-compatible 16-byte records in
// page-locked memory, so the buffer can go straight into plane_ransac_set_cloud_async (chunked, overlapped upload).
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/plane_ransac.h"

namespace pr {
int pcd_fail(int code, const char* fmt, ...);  // pr_api.cpp: sets plane_ransac_last_error
void* plain_alloc(size_t bytes);               // pr_api.cpp: pageable fallback that plane_ransac_host_free understands
}

namespace {

struct Field {
  std::string name;
  int size = 4, count = 1;
  char type = 'F';
  size_t offset = 0;  // byte offset inside one record (binary) / column (ascii) / block start (compressed: elements)
};

double read_scalar(const unsigned char* p, char type, int size) {
  switch (type) {
    case 'F':
      if (size == 4) { float v; std::memcpy(&v, p, 4); return v; }
      if (size == 8) { double v; std::memcpy(&v, p, 8); return v; }
      break;
    case 'U':
      if (size == 1) return *p;
      if (size == 2) { uint16_t v; std::memcpy(&v, p, 2); return v; }
      if (size == 4) { uint32_t v; std::memcpy(&v, p, 4); return v; }
      if (size == 8) { uint64_t v; std::memcpy(&v, p, 8); return (double)v; }
      break;
    case 'I':
      if (size == 1) return (int8_t)*p;
      if (size == 2) { int16_t v; std::memcpy(&v, p, 2); return v; }
      if (size == 4) { int32_t v; std::memcpy(&v, p, 4); return v; }
      if (size == 8) { int64_t v; std::memcpy(&v, p, 8); return (double)v; }
      break;
  }
  return 0.0;
}

// liblzf's decompressor (the format pcl::lzfDecompress reads): control byte < 32 = literal run, else back reference
bool lzf_decompress(const unsigned char* in, size_t in_len, unsigned char* out, size_t out_len) {
  const unsigned char* ip = in;
  const unsigned char* const in_end = in + in_len;
  unsigned char* op = out;
  unsigned char* const out_end = out + out_len;
  while (ip < in_end) {
    unsigned ctrl = *ip++;
    if (ctrl < 32) {
      ++ctrl;
      if (op + ctrl > out_end || ip + ctrl > in_end) return false;
      std::memcpy(op, ip, ctrl);
      op += ctrl;
      ip += ctrl;
    } else {
      unsigned len = ctrl >> 5;
      if (ip >= in_end) return false;
      size_t back = ((size_t)(ctrl & 0x1f) << 8) + 1;
      if (len == 7) {
        len += *ip++;
        if (ip >= in_end) return false;
      }
      back += *ip++;
      len += 2;
      if (op + len > out_end || (size_t)(op - out) < back) return false;
      const unsigned char* ref = op - back;
      for (unsigned i = 0; i < len; ++i) *op++ = *ref++;  // byte by byte: the ranges may overlap
    }
  }
  return op == out_end;
}

}  // namespace

extern "C" int plane_ransac_load_pcd(const char* path, pr_point** points, size_t* n_points) {
  if (!path || !points || !n_points) return pr::pcd_fail(PR_ERR_INVALID, "null argument");
  *points = nullptr;
  *n_points = 0;
  FILE* f = std::fopen(path, "rb");
  if (!f) return pr::pcd_fail(PR_ERR_INVALID, "cannot open %s: %s", path, std::strerror(errno));
  std::vector<unsigned char> raw;
  {
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    raw.resize(sz > 0 ? (size_t)sz : 0);
    const size_t got = raw.empty() ? 0 : std::fread(raw.data(), 1, raw.size(), f);
    std::fclose(f);
    if (got != raw.size()) return pr::pcd_fail(PR_ERR_INVALID, "short read on %s", path);
  }
  // ---- header: "KEY v v v" lines up to and including DATA
  std::vector<std::string> names, sizes, types, counts;
  long long width = -1, height = 1, n_decl = -1;
  std::string data_kind;
  size_t pos = 0;
  while (pos < raw.size()) {
    size_t end = pos;
    while (end < raw.size() && raw[end] != '\n') ++end;
    std::string line(reinterpret_cast<const char*>(raw.data()) + pos, end - pos);
    pos = end < raw.size() ? end + 1 : end;
    std::vector<std::string> tok;
    size_t i = 0;
    while (i < line.size()) {
      while (i < line.size() && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r')) ++i;
      size_t j = i;
      while (j < line.size() && line[j] != ' ' && line[j] != '\t' && line[j] != '\r') ++j;
      if (j > i) tok.emplace_back(line.substr(i, j - i));
      i = j;
    }
    if (tok.empty() || tok[0][0] == '#') continue;
    std::string key = tok[0];
    for (char& ch : key) ch = (char)std::toupper((unsigned char)ch);
    std::vector<std::string> vals(tok.begin() + 1, tok.end());
    if (key == "FIELDS" || key == "COLUMNS") names = vals;
    else if (key == "SIZE") sizes = vals;
    else if (key == "TYPE") types = vals;
    else if (key == "COUNT") counts = vals;
    else if (key == "WIDTH" && !vals.empty()) width = std::atoll(vals[0].c_str());
    else if (key == "HEIGHT" && !vals.empty()) height = std::atoll(vals[0].c_str());
    else if (key == "POINTS" && !vals.empty()) n_decl = std::atoll(vals[0].c_str());
    else if (key == "DATA") {
      if (!vals.empty()) data_kind = vals[0];
      for (char& ch : data_kind) ch = (char)std::tolower((unsigned char)ch);
      break;
    }
  }
  if (data_kind.empty()) return pr::pcd_fail(PR_ERR_INVALID, "%s: no DATA line (not a PCD file?)", path);
  if (names.empty() || sizes.size() != names.size() || types.size() != names.size())
    return pr::pcd_fail(PR_ERR_INVALID, "%s: FIELDS / SIZE / TYPE do not match", path);
  std::vector<Field> fields(names.size());
  int xyz[3] = {-1, -1, -1};
  size_t rec_bytes = 0, rec_cols = 0;
  for (size_t k = 0; k < names.size(); ++k) {
    Field& fd = fields[k];
    fd.name = names[k];
    fd.size = std::atoi(sizes[k].c_str());
    fd.type = (char)std::toupper((unsigned char)types[k][0]);
    fd.count = k < counts.size() ? std::atoi(counts[k].c_str()) : 1;
    if (fd.count < 0 || (fd.size != 1 && fd.size != 2 && fd.size != 4 && fd.size != 8) || (fd.type != 'F' && fd.type != 'U' && fd.type != 'I') ||
        (fd.type == 'F' && fd.size < 4))
      return pr::pcd_fail(PR_ERR_INVALID, "%s: unsupported field %s (%c%d)", path, fd.name.c_str(), fd.type, fd.size);
    if (fd.name == "x") xyz[0] = (int)k;
    if (fd.name == "y") xyz[1] = (int)k;
    if (fd.name == "z") xyz[2] = (int)k;
    rec_bytes += (size_t)fd.size * (size_t)fd.count;
    rec_cols += (size_t)fd.count;
  }
  if (xyz[0] < 0 || xyz[1] < 0 || xyz[2] < 0) return pr::pcd_fail(PR_ERR_INVALID, "%s: no x y z fields", path);
  long long n_ll = n_decl >= 0 ? n_decl : (width >= 0 ? width * height : -1);
  if (n_ll < 0 || n_ll > (long long)INT32_MAX - 4096) return pr::pcd_fail(PR_ERR_INVALID, "%s: bad point count", path);
  const size_t n = (size_t)n_ll;
  void* mem = nullptr;
  if (plane_ransac_host_alloc((n ? n : 1) * sizeof(pr_point), &mem) != PR_OK) {
    mem = pr::plain_alloc((n ? n : 1) * sizeof(pr_point));  // no CUDA device to page-lock for: pageable memory
    if (!mem) return pr::pcd_fail(PR_ERR_OOM, "%s: out of host memory for %zu points", path, n);
  }
  pr_point* out = static_cast<pr_point*>(mem);
  auto bail = [&](const char* what) {
    plane_ransac_host_free(out);
    return pr::pcd_fail(PR_ERR_INVALID, "%s: %s", path, what);
  };
  if (data_kind == "ascii") {
    size_t col_of[3] = {0, 0, 0}, c = 0;
    for (size_t k = 0; k < fields.size(); ++k) {
      for (int a = 0; a < 3; ++a)
        if (xyz[a] == (int)k) col_of[a] = c;
      c += (size_t)fields[k].count;
    }
    raw.push_back(0);  // terminator for strtof
    const char* p = reinterpret_cast<const char*>(raw.data()) + pos;
    for (size_t i = 0; i < n; ++i) {
      float v[3] = {0, 0, 0};
      for (size_t col = 0; col < rec_cols; ++col) {
        while (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n') ++p;
        if (!*p) return bail("fewer values than POINTS x fields");
        char* e = nullptr;
        const float val = std::strtof(p, &e);  // correctly rounded, "nan" / "inf" included — as the stream extraction PCL uses
        if (e == p) return bail("unreadable value in the ascii body");
        for (int a = 0; a < 3; ++a)
          if (col == col_of[a]) v[a] = val;
        p = e;
      }
      out[i].x = v[0]; out[i].y = v[1]; out[i].z = v[2]; out[i].w = 1.0f;
    }
  } else if (data_kind == "binary") {
    size_t off = 0;
    for (Field& fd : fields) { fd.offset = off; off += (size_t)fd.size * (size_t)fd.count; }
    if (raw.size() < pos + n * rec_bytes) return bail("binary body shorter than POINTS x record size");
    const unsigned char* body = raw.data() + pos;
    for (size_t i = 0; i < n; ++i) {
      const unsigned char* r = body + i * rec_bytes;
      out[i].x = (float)read_scalar(r + fields[xyz[0]].offset, fields[xyz[0]].type, fields[xyz[0]].size);
      out[i].y = (float)read_scalar(r + fields[xyz[1]].offset, fields[xyz[1]].type, fields[xyz[1]].size);
      out[i].z = (float)read_scalar(r + fields[xyz[2]].offset, fields[xyz[2]].type, fields[xyz[2]].size);
      out[i].w = 1.0f;
    }
  } else if (data_kind == "binary_compressed") {
    if (raw.size() < pos + 8) return bail("truncated compressed header");
    uint32_t comp = 0, uncomp = 0;
    std::memcpy(&comp, raw.data() + pos, 4);
    std::memcpy(&uncomp, raw.data() + pos + 4, 4);
    if (raw.size() < pos + 8 + comp || (size_t)uncomp != n * rec_bytes) return bail("compressed sizes do not match the header");
    std::vector<unsigned char> buf(uncomp ? uncomp : 1);
    if (uncomp && !lzf_decompress(raw.data() + pos + 8, comp, buf.data(), uncomp)) return bail("LZF stream is corrupt");
    size_t off = 0;  // field-major: all of field 0 for every point, then field 1, ...
    for (Field& fd : fields) { fd.offset = off; off += (size_t)fd.size * (size_t)fd.count * n; }
    for (size_t i = 0; i < n; ++i) {
      float v[3];
      for (int a = 0; a < 3; ++a) {
        const Field& fd = fields[xyz[a]];
        v[a] = (float)read_scalar(buf.data() + fd.offset + i * (size_t)fd.size * (size_t)fd.count, fd.type, fd.size);
      }
      out[i].x = v[0]; out[i].y = v[1]; out[i].z = v[2]; out[i].w = 1.0f;
    }
  } else {
    return bail("unknown DATA kind");
  }
  *points = out;
  *n_points = n;
  return PR_OK;
}
