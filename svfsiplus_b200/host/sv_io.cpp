// sv_io.cpp — VTK-free mesh / result / restart / history I/O behind include/svb200_io.h.
// Host only (C++17 + zlib).  What each part replaces in the reference is cited in the header; the formats themselves are
//   * VTK XML (Kitware, "VTK File Formats", XML section): <VTKFile type byte_order header_type compressor> / <Piece> /
//     <Points> <Cells|Polys> <PointData> <CellData> with <DataArray type Name NumberOfComponents format [offset]>;
//     binary blocks are  [nbytes][data]  or, compressed,  [nblocks][usize][psize][csize_1..n][zlib blocks],  header words
//     UInt32 or UInt64, inline / appended-base64 blocks base64-encode header and payload SEPARATELY, appended-raw blocks
//     start at `offset` bytes after the '_' of <AppendedData encoding="raw">;
//   * the restart record of output::write_restart and the history line of output::output_result (plain streams).
#include "svb200_io.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

enum class VT { I8, U8, I16, U16, I32, U32, I64, U64, F32, F64 };

struct TypeInfo { const char* name; VT t; int size; bool integer; };
const TypeInfo kTypes[] = {
  {"Int8", VT::I8, 1, true}, {"UInt8", VT::U8, 1, true}, {"Int16", VT::I16, 2, true}, {"UInt16", VT::U16, 2, true},
  {"Int32", VT::I32, 4, true}, {"UInt32", VT::U32, 4, true}, {"Int64", VT::I64, 8, true}, {"UInt64", VT::U64, 8, true},
  {"Float32", VT::F32, 4, false}, {"Float64", VT::F64, 8, false},
};

const TypeInfo& type_by_name(const std::string& n)
{
  for (auto& t : kTypes) if (n == t.name) return t;
  throw std::runtime_error("unknown DataArray type '" + n + "'");
}
const TypeInfo& type_info(VT v)
{
  for (auto& t : kTypes) if (t.t == v) return t;
  throw std::runtime_error("bad type");
}

struct DataArray {
  std::string name;
  VT type = VT::F64;
  int ncomp = 1;
  std::vector<unsigned char> bytes;          // native little-endian values, tuple-major
  size_t count() const { return bytes.size() / type_info(type).size; }
  template <class T> T at(size_t i) const
  {
    const unsigned char* p = bytes.data() + i*type_info(type).size;
    switch (type) {
      case VT::I8:  { int8_t v;   memcpy(&v, p, 1); return T(v); }
      case VT::U8:  { uint8_t v;  memcpy(&v, p, 1); return T(v); }
      case VT::I16: { int16_t v;  memcpy(&v, p, 2); return T(v); }
      case VT::U16: { uint16_t v; memcpy(&v, p, 2); return T(v); }
      case VT::I32: { int32_t v;  memcpy(&v, p, 4); return T(v); }
      case VT::U32: { uint32_t v; memcpy(&v, p, 4); return T(v); }
      case VT::I64: { int64_t v;  memcpy(&v, p, 8); return T(v); }
      case VT::U64: { uint64_t v; memcpy(&v, p, 8); return T(v); }
      case VT::F32: { float v;    memcpy(&v, p, 4); return T(v); }
      default:      { double v;   memcpy(&v, p, 8); return T(v); }
    }
  }
};

template <class T> DataArray make_array(const std::string& name, VT t, int ncomp, const T* data, size_t n)
{
  DataArray a;
  a.name = name; a.type = t; a.ncomp = ncomp;
  a.bytes.resize(n*sizeof(T));
  if (n) memcpy(a.bytes.data(), data, n*sizeof(T));
  return a;
}

} // namespace

struct b200io_vtk {
  bool polydata = false;
  int nNo = 0, nEl = 0;
  std::vector<double> x;                     // 3 x nNo
  std::vector<int64_t> conn, offsets;        // offsets: exclusive end of each cell (VTK convention)
  std::vector<unsigned char> types;
  std::vector<DataArray> pdata, cdata;
  std::vector<DataArray>& arrays(int where) { return where == B200IO_CELL_DATA ? cdata : pdata; }
  const std::vector<DataArray>& arrays(int where) const { return where == B200IO_CELL_DATA ? cdata : pdata; }
};

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// base64
// ---------------------------------------------------------------------------------------------------------------------
const char kB64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";

void b64_encode(const unsigned char* p, size_t n, std::string& out)
{
  size_t i = 0;
  for (; i + 2 < n; i += 3) {
    const unsigned v = (unsigned(p[i]) << 16) | (unsigned(p[i+1]) << 8) | p[i+2];
    out.push_back(kB64[(v >> 18) & 63]); out.push_back(kB64[(v >> 12) & 63]);
    out.push_back(kB64[(v >> 6) & 63]);  out.push_back(kB64[v & 63]);
  }
  if (i + 1 == n) {
    const unsigned v = unsigned(p[i]) << 16;
    out.push_back(kB64[(v >> 18) & 63]); out.push_back(kB64[(v >> 12) & 63]); out += "==";
  } else if (i + 2 == n) {
    const unsigned v = (unsigned(p[i]) << 16) | (unsigned(p[i+1]) << 8);
    out.push_back(kB64[(v >> 18) & 63]); out.push_back(kB64[(v >> 12) & 63]); out.push_back(kB64[(v >> 6) & 63]);
    out.push_back('=');
  }
}

// Incremental decoder over a character range: yields bytes group by group.  A padded group ends one separately encoded
// piece (VTK encodes the block header and the payload separately); the next group simply starts the next piece, so the
// concatenation of everything decoded is header bytes followed by payload bytes in both conventions.
struct B64Reader {
  const char* p; const char* end;
  signed char lut[256];
  B64Reader(const char* b, const char* e) : p(b), end(e)
  {
    memset(lut, -1, sizeof(lut));
    for (int i = 0; i < 64; i++) lut[(unsigned char)kB64[i]] = (signed char)i;
  }
  // append at least `need` more bytes to out (or stop at the end of the text); returns false when the text ran out first
  bool read(std::vector<unsigned char>& out, size_t need)
  {
    const size_t target = out.size() + need;
    while (out.size() < target) {
      int q[4], k = 0, pad = 0;
      while (k < 4 && p < end) {
        const unsigned char c = (unsigned char)*p++;
        if (c == '=') { q[k++] = 0; pad++; }
        else if (lut[c] >= 0) { if (pad) throw std::runtime_error("base64: data after padding inside a group"); q[k++] = lut[c]; }
        else if (c == ' ' || c == '\n' || c == '\r' || c == '\t') continue;
        else throw std::runtime_error("base64: invalid character");
      }
      if (k == 0) return false;
      if (k < 4) throw std::runtime_error("base64: truncated group");
      const unsigned v = (unsigned(q[0]) << 18) | (unsigned(q[1]) << 12) | (unsigned(q[2]) << 6) | unsigned(q[3]);
      out.push_back((unsigned char)(v >> 16));
      if (pad < 2) out.push_back((unsigned char)(v >> 8));
      if (pad < 1) out.push_back((unsigned char)v);
    }
    return true;
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// binary blocks
// ---------------------------------------------------------------------------------------------------------------------
uint64_t header_word(const unsigned char* p, int hsize)
{
  if (hsize == 4) { uint32_t v; memcpy(&v, p, 4); return v; }
  uint64_t v; memcpy(&v, p, 8); return v;
}
void put_word(std::vector<unsigned char>& out, uint64_t v, int hsize)
{
  if (hsize == 4) { uint32_t w = uint32_t(v); out.insert(out.end(), (unsigned char*)&w, (unsigned char*)&w + 4); }
  else out.insert(out.end(), (unsigned char*)&v, (unsigned char*)&v + 8);
}

void inflate_block(const unsigned char* src, size_t csize, unsigned char* dst, size_t usize)
{
  uLongf dl = uLongf(usize);
  const int rc = uncompress(dst, &dl, src, uLong(csize));
  if (rc != Z_OK || dl != usize) throw std::runtime_error("zlib: cannot inflate a data block");
}

// A "source" hands out bytes sequentially: raw memory or a base64 text.
struct ByteSource {
  const unsigned char* raw = nullptr; size_t raw_n = 0, raw_pos = 0;
  B64Reader* b64 = nullptr; std::vector<unsigned char> buf; size_t buf_pos = 0;
  // upper bound of the bytes this source can still deliver (header fields of a corrupt file must not drive an allocation)
  size_t available() const
  {
    if (raw) return raw_n - raw_pos;
    return (buf.size() - buf_pos) + (size_t(b64->end - b64->p)/4 + 1)*3;
  }
  void take(size_t n, std::vector<unsigned char>& out)
  {
    if (n > available()) throw std::runtime_error(raw ? "appended data: block runs past the end of the file"
                                                      : "base64 data: block is shorter than its header says");
    out.resize(n);
    if (raw) {
      if (raw_pos + n > raw_n) throw std::runtime_error("appended data: block runs past the end of the file");
      if (n) memcpy(out.data(), raw + raw_pos, n);
      raw_pos += n;
    } else {
      if (buf.size() - buf_pos < n && !b64->read(buf, n - (buf.size() - buf_pos)))
        throw std::runtime_error("base64 data: block is shorter than its header says");
      if (n) memcpy(out.data(), buf.data() + buf_pos, n);
      buf_pos += n;
    }
  }
};

std::vector<unsigned char> read_block(ByteSource& src, int hsize, bool compressed)
{
  std::vector<unsigned char> h, out;
  if (!compressed) {
    src.take(hsize, h);
    src.take(size_t(header_word(h.data(), hsize)), out);
    return out;
  }
  src.take(3*size_t(hsize), h);
  const uint64_t nb = header_word(h.data(), hsize), us = header_word(h.data() + hsize, hsize), ps = header_word(h.data() + 2*hsize, hsize);
  if (nb == 0) return out;
  // header sanity before anything is sized from it: the last block is a partial block (ps <= us), the block table must fit in
  // what is left of the input, and zlib cannot expand a block by more than ~1032 : 1
  if (us == 0 || ps > us) throw std::runtime_error("compressed data: inconsistent block header (partial block larger than the block size)");
  if (nb > src.available()/size_t(hsize)) throw std::runtime_error("compressed data: block table runs past the end of the input");
  std::vector<unsigned char> cs;
  src.take(size_t(nb)*hsize, cs);
  const size_t total = size_t(nb - 1)*us + (ps ? ps : us);
  if (total/1100 > src.available()) throw std::runtime_error("compressed data: header announces more data than the input can hold");
  out.resize(total);
  size_t off = 0;
  std::vector<unsigned char> blk;
  for (uint64_t b = 0; b < nb; b++) {
    const size_t c = size_t(header_word(cs.data() + b*hsize, hsize));
    const size_t u = (b + 1 == nb && ps) ? size_t(ps) : size_t(us);
    src.take(c, blk);
    inflate_block(blk.data(), c, out.data() + off, u);
    off += u;
  }
  return out;
}

void write_block(const unsigned char* p, size_t n, int hsize, bool compress, std::vector<unsigned char>& header, std::vector<unsigned char>& payload)
{
  header.clear(); payload.clear();
  if (!compress) {
    put_word(header, n, hsize);
    payload.assign(p, p + n);
    return;
  }
  const size_t bs = 32768;                   // vtkXMLWriter's default block size
  const size_t nb = n ? (n + bs - 1)/bs : 0;
  const size_t last = n ? n - (nb - 1)*bs : 0;
  put_word(header, nb, hsize);
  put_word(header, bs, hsize);
  put_word(header, (last == bs) ? 0 : last, hsize);
  std::vector<unsigned char> tmp(compressBound(uLong(bs)));
  for (size_t b = 0; b < nb; b++) {
    const size_t u = (b + 1 == nb) ? last : bs;
    uLongf cl = uLongf(tmp.size());
    if (compress2(tmp.data(), &cl, p + b*bs, uLong(u), Z_DEFAULT_COMPRESSION) != Z_OK) throw std::runtime_error("zlib: compress failed");
    put_word(header, cl, hsize);
    payload.insert(payload.end(), tmp.data(), tmp.data() + cl);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// a small XML scanner: elements with attributes and character data, enough for VTK XML files.  The appended-data section
// (raw bytes, not XML) is cut out before scanning.
// ---------------------------------------------------------------------------------------------------------------------
struct Node {
  std::string tag;
  std::map<std::string, std::string> attr;
  size_t text_begin = 0, text_end = 0;       // character data up to the first child / the end tag
  std::vector<Node> kids;
  const Node* child(const std::string& t) const { for (auto& k : kids) if (k.tag == t) return &k; return nullptr; }
  std::string get(const std::string& k, const std::string& dflt = "") const { auto it = attr.find(k); return it == attr.end() ? dflt : it->second; }
};

struct Scanner {
  const std::string& s; size_t pos = 0;
  explicit Scanner(const std::string& str) : s(str) {}
  void skip_ws() { while (pos < s.size() && isspace((unsigned char)s[pos])) pos++; }
  bool starts(const char* lit) const { return s.compare(pos, strlen(lit), lit) == 0; }
  void skip_misc()
  {
    for (;;) {
      skip_ws();
      if (starts("<?")) { pos = s.find("?>", pos); if (pos == std::string::npos) throw std::runtime_error("xml: unterminated declaration"); pos += 2; }
      else if (starts("<!--")) { pos = s.find("-->", pos); if (pos == std::string::npos) throw std::runtime_error("xml: unterminated comment"); pos += 3; }
      else if (starts("<!")) { pos = s.find('>', pos); if (pos == std::string::npos) throw std::runtime_error("xml: unterminated doctype"); pos += 1; }
      else return;
    }
  }
  static std::string unescape(const std::string& v)
  {
    if (v.find('&') == std::string::npos) return v;
    std::string o;
    for (size_t i = 0; i < v.size(); i++) {
      if (v[i] != '&') { o.push_back(v[i]); continue; }
      if (v.compare(i, 4, "&lt;") == 0) { o.push_back('<'); i += 3; }
      else if (v.compare(i, 4, "&gt;") == 0) { o.push_back('>'); i += 3; }
      else if (v.compare(i, 5, "&amp;") == 0) { o.push_back('&'); i += 4; }
      else if (v.compare(i, 6, "&quot;") == 0) { o.push_back('"'); i += 5; }
      else if (v.compare(i, 6, "&apos;") == 0) { o.push_back('\''); i += 5; }
      else o.push_back('&');
    }
    return o;
  }
  Node element()
  {
    skip_misc();
    if (pos >= s.size() || s[pos] != '<') throw std::runtime_error("xml: expected an element");
    pos++;
    Node n;
    size_t b = pos;
    while (pos < s.size() && !isspace((unsigned char)s[pos]) && s[pos] != '>' && s[pos] != '/') pos++;
    n.tag = s.substr(b, pos - b);
    for (;;) {
      skip_ws();
      if (pos >= s.size()) throw std::runtime_error("xml: unterminated tag <" + n.tag);
      if (s[pos] == '/') { pos++; if (pos >= s.size() || s[pos] != '>') throw std::runtime_error("xml: bad empty-element tag"); pos++; return n; }
      if (s[pos] == '>') { pos++; break; }
      b = pos;
      while (pos < s.size() && s[pos] != '=' && !isspace((unsigned char)s[pos])) pos++;
      const std::string key = s.substr(b, pos - b);
      skip_ws();
      if (pos >= s.size() || s[pos] != '=') throw std::runtime_error("xml: attribute without a value in <" + n.tag);
      pos++; skip_ws();
      const char q = s[pos];
      if (q != '"' && q != '\'') throw std::runtime_error("xml: unquoted attribute in <" + n.tag);
      const size_t e = s.find(q, pos + 1);
      if (e == std::string::npos) throw std::runtime_error("xml: unterminated attribute in <" + n.tag);
      n.attr[key] = unescape(s.substr(pos + 1, e - pos - 1));
      pos = e + 1;
    }
    n.text_begin = pos;
    bool have_text_end = false;
    for (;;) {
      const size_t lt = s.find('<', pos);
      if (lt == std::string::npos) throw std::runtime_error("xml: missing </" + n.tag + ">");
      if (!have_text_end) { n.text_end = lt; have_text_end = true; }
      pos = lt;
      if (starts("</")) {
        const size_t e = s.find('>', pos);
        if (e == std::string::npos) throw std::runtime_error("xml: unterminated end tag");
        std::string t = s.substr(pos + 2, e - pos - 2);
        while (!t.empty() && isspace((unsigned char)t.back())) t.pop_back();
        if (t != n.tag) throw std::runtime_error("xml: </" + t + "> closes <" + n.tag + ">");
        pos = e + 1;
        return n;
      }
      if (starts("<!--")) { pos = s.find("-->", pos); if (pos == std::string::npos) throw std::runtime_error("xml: unterminated comment"); pos += 3; continue; }
      n.kids.push_back(element());
    }
  }
};

struct FileCtx {
  std::string xml;                           // the file without the appended payload
  std::vector<unsigned char> appended;       // bytes after '_' (raw) or the base64 text
  bool appended_raw = false, have_appended = false;
  int hsize = 4;
  bool compressed = false;
};

DataArray decode_array(const Node& da, const FileCtx& f)
{
  DataArray a;
  a.name = da.get("Name");
  const TypeInfo& ti = type_by_name(da.get("type", "Float64"));
  a.type = ti.t;
  a.ncomp = std::max(1, atoi(da.get("NumberOfComponents", "1").c_str()));
  const std::string fmt = da.get("format", "ascii");
  if (fmt == "ascii") {
    const char* p = f.xml.c_str() + da.text_begin;
    const char* e = f.xml.c_str() + da.text_end;
    std::vector<unsigned char>& out = a.bytes;
    while (p < e) {
      while (p < e && isspace((unsigned char)*p)) p++;
      if (p >= e) break;
      char* q = nullptr;
      if (ti.integer) {
        if (ti.t == VT::U64) { const uint64_t v = strtoull(p, &q, 10); if (q == p) throw std::runtime_error("ascii DataArray '" + a.name + "': not a number"); out.insert(out.end(), (unsigned char*)&v, (unsigned char*)&v + 8); }
        else {
          const long long v = strtoll(p, &q, 10);
          if (q == p) throw std::runtime_error("ascii DataArray '" + a.name + "': not a number");
          switch (ti.t) {
            case VT::I8:  { int8_t w = int8_t(v);   out.insert(out.end(), (unsigned char*)&w, (unsigned char*)&w + 1); break; }
            case VT::U8:  { uint8_t w = uint8_t(v); out.insert(out.end(), (unsigned char*)&w, (unsigned char*)&w + 1); break; }
            case VT::I16: { int16_t w = int16_t(v); out.insert(out.end(), (unsigned char*)&w, (unsigned char*)&w + 2); break; }
            case VT::U16: { uint16_t w = uint16_t(v); out.insert(out.end(), (unsigned char*)&w, (unsigned char*)&w + 2); break; }
            case VT::I32: { int32_t w = int32_t(v); out.insert(out.end(), (unsigned char*)&w, (unsigned char*)&w + 4); break; }
            case VT::U32: { uint32_t w = uint32_t(v); out.insert(out.end(), (unsigned char*)&w, (unsigned char*)&w + 4); break; }
            default:      { int64_t w = int64_t(v); out.insert(out.end(), (unsigned char*)&w, (unsigned char*)&w + 8); break; }
          }
        }
      } else {
        const double v = strtod(p, &q);
        if (q == p) throw std::runtime_error("ascii DataArray '" + a.name + "': not a number");
        if (ti.t == VT::F32) { const float w = float(v); out.insert(out.end(), (unsigned char*)&w, (unsigned char*)&w + 4); }
        else out.insert(out.end(), (unsigned char*)&v, (unsigned char*)&v + 8);
      }
      p = q;
    }
  } else if (fmt == "binary") {
    B64Reader r(f.xml.c_str() + da.text_begin, f.xml.c_str() + da.text_end);
    ByteSource src; src.b64 = &r;
    a.bytes = read_block(src, f.hsize, f.compressed);
  } else if (fmt == "appended") {
    if (!f.have_appended) throw std::runtime_error("DataArray '" + a.name + "' is appended but the file has no <AppendedData>");
    const size_t off = size_t(strtoull(da.get("offset", "0").c_str(), nullptr, 10));
    if (off > f.appended.size()) throw std::runtime_error("DataArray '" + a.name + "': offset past the end of the appended data");
    if (f.appended_raw) {
      ByteSource src; src.raw = f.appended.data(); src.raw_n = f.appended.size(); src.raw_pos = off;
      a.bytes = read_block(src, f.hsize, f.compressed);
    } else {
      B64Reader r((const char*)f.appended.data() + off, (const char*)f.appended.data() + f.appended.size());
      ByteSource src; src.b64 = &r;
      a.bytes = read_block(src, f.hsize, f.compressed);
    }
  } else throw std::runtime_error("DataArray '" + a.name + "': unknown format '" + fmt + "'");
  if (a.bytes.size() % ti.size) throw std::runtime_error("DataArray '" + a.name + "': byte count is not a multiple of the type size");
  return a;
}

std::vector<int64_t> as_i64(const DataArray& a)
{
  std::vector<int64_t> v(a.count());
  for (size_t i = 0; i < v.size(); i++) v[i] = a.at<int64_t>(i);
  return v;
}

const Node* named_array(const Node& sec, const char* name)
{
  for (auto& k : sec.kids) if (k.tag == "DataArray" && k.get("Name") == name) return &k;
  return nullptr;
}

b200io_vtk* read_file(const char* path)
{
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error(std::string("cannot open '") + path + "'");
  std::string all((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  FileCtx f;
  const size_t ap = all.find("<AppendedData");
  if (ap != std::string::npos) {
    const size_t gt = all.find('>', ap);
    if (gt == std::string::npos) throw std::runtime_error("unterminated <AppendedData>");
    const std::string tag = all.substr(ap, gt - ap);
    f.appended_raw = tag.find("raw") != std::string::npos;
    const size_t us = all.find('_', gt);
    const size_t close = all.rfind("</AppendedData>");
    if (us == std::string::npos || close == std::string::npos || close < us) throw std::runtime_error("malformed <AppendedData> section");
    size_t e = close;
    if (!f.appended_raw) while (e > us + 1 && isspace((unsigned char)all[e-1])) e--;
    f.appended.assign(all.begin() + us + 1, all.begin() + e);
    f.have_appended = true;
    f.xml = all.substr(0, ap) + all.substr(close + strlen("</AppendedData>"));
  } else f.xml.swap(all);

  Scanner sc(f.xml);
  const Node root = sc.element();
  if (root.tag != "VTKFile") throw std::runtime_error("not a VTK XML file (root element <" + root.tag + ">)");
  if (root.get("byte_order", "LittleEndian") != "LittleEndian") throw std::runtime_error("BigEndian VTK files are not supported");
  const std::string ht = root.get("header_type", "UInt32");
  if (ht != "UInt32" && ht != "UInt64") throw std::runtime_error("unknown header_type '" + ht + "'");
  f.hsize = (ht == "UInt64") ? 8 : 4;
  const std::string comp = root.get("compressor");
  if (!comp.empty() && comp != "vtkZLibDataCompressor") throw std::runtime_error("compressor '" + comp + "' is not supported (zlib only)");
  f.compressed = !comp.empty();
  const std::string type = root.get("type");
  if (type != "UnstructuredGrid" && type != "PolyData") throw std::runtime_error("VTK file type '" + type + "' is not supported");
  const Node* grid = root.child(type);
  if (!grid) throw std::runtime_error("missing <" + type + "> element");
  int npieces = 0;
  for (auto& k : grid->kids) if (k.tag == "Piece") npieces++;
  if (npieces != 1) throw std::runtime_error("expected exactly one <Piece> (parallel pieces are not supported)");
  const Node& piece = *grid->child("Piece");

  std::unique_ptr<b200io_vtk> h(new b200io_vtk);
  h->polydata = (type == "PolyData");
  h->nNo = atoi(piece.get("NumberOfPoints", "0").c_str());
  const Node* pts = piece.child("Points");
  if (h->nNo > 0) {
    if (!pts || pts->kids.empty()) throw std::runtime_error("missing <Points>");
    const DataArray pa = decode_array(pts->kids[0], f);
    if (pa.ncomp != 3 || pa.count() != size_t(3)*h->nNo) throw std::runtime_error("<Points> must hold 3 x NumberOfPoints values");
    h->x.resize(size_t(3)*h->nNo);
    for (size_t i = 0; i < h->x.size(); i++) h->x[i] = pa.at<double>(i);
  }
  const Node* cells = piece.child(h->polydata ? "Polys" : "Cells");
  h->nEl = atoi(piece.get(h->polydata ? "NumberOfPolys" : "NumberOfCells", "0").c_str());
  if (h->polydata && h->nEl == 0 && atoi(piece.get("NumberOfLines", "0").c_str()) > 0) {     // 2-D boundary "faces" are poly-lines
    cells = piece.child("Lines");
    h->nEl = atoi(piece.get("NumberOfLines", "0").c_str());
  }
  if (h->nEl > 0) {
    if (!cells) throw std::runtime_error("missing cell section");
    const Node* c = named_array(*cells, "connectivity");
    const Node* o = named_array(*cells, "offsets");
    if (!c || !o) throw std::runtime_error("cell section needs 'connectivity' and 'offsets'");
    h->conn = as_i64(decode_array(*c, f));
    h->offsets = as_i64(decode_array(*o, f));
    if (h->offsets.size() != size_t(h->nEl)) throw std::runtime_error("'offsets' must hold one entry per cell");
    int64_t prev = 0;
    for (auto e : h->offsets) { if (e < prev || size_t(e) > h->conn.size()) throw std::runtime_error("'offsets' is not a valid prefix sum of 'connectivity'"); prev = e; }
    for (auto n : h->conn) if (n < 0 || n >= h->nNo) throw std::runtime_error("'connectivity' refers to a point that does not exist");
    h->types.resize(h->nEl);
    const Node* t = h->polydata ? nullptr : named_array(*cells, "types");
    if (t) {
      const DataArray ta = decode_array(*t, f);
      if (ta.count() != size_t(h->nEl)) throw std::runtime_error("'types' must hold one entry per cell");
      for (int e = 0; e < h->nEl; e++) h->types[e] = ta.at<unsigned char>(e);
    } else {
      int64_t b = 0;
      for (int e = 0; e < h->nEl; e++) { const int64_t n = h->offsets[e] - b; b = h->offsets[e]; h->types[e] = (n == 2) ? 3 : (n == 3) ? 5 : (n == 4) ? 9 : 7; }
    }
  }
  for (int where = 0; where < 2; where++) {
    const Node* sec = piece.child(where ? "CellData" : "PointData");
    if (!sec) continue;
    const size_t ntup = where ? h->nEl : h->nNo;
    for (auto& k : sec->kids) {
      if (k.tag != "DataArray") continue;
      DataArray a = decode_array(k, f);
      if (a.count() != ntup*a.ncomp) throw std::runtime_error("DataArray '" + a.name + "' does not hold NumberOfComponents x tuples values");
      h->arrays(where).push_back(std::move(a));
    }
  }
  return h.release();
}

// ---------------------------------------------------------------------------------------------------------------------
// writer
// ---------------------------------------------------------------------------------------------------------------------
struct Writer {
  int mode, hsize; bool compress;
  std::string xml;
  std::vector<unsigned char> app_raw; std::string app_b64;
  void array(const DataArray& a, const char* indent)
  {
    const TypeInfo& ti = type_info(a.type);
    char buf[512];
    std::string esc;                        // attribute value: the five XML entities
    for (char ch : a.name) {
      switch (ch) {
        case '&': esc += "&amp;"; break;
        case '<': esc += "&lt;"; break;
        case '>': esc += "&gt;"; break;
        case '"': esc += "&quot;"; break;
        case '\'': esc += "&apos;"; break;
        default: esc.push_back(ch);
      }
    }
    xml += std::string(indent) + "<DataArray type=\"" + ti.name + "\" Name=\"" + esc + "\" NumberOfComponents=\"" + std::to_string(a.ncomp) +
           "\" format=\"" + (mode == B200IO_ASCII ? "ascii" : mode == B200IO_BINARY ? "binary" : "appended") + "\"";
    if (mode == B200IO_ASCII) {
      xml += ">\n";
      const size_t n = a.count();
      std::string line;
      for (size_t i = 0; i < n; i++) {
        if (ti.integer) { if (a.type == VT::U64) snprintf(buf, sizeof(buf), "%llu", (unsigned long long)a.at<uint64_t>(i)); else snprintf(buf, sizeof(buf), "%lld", (long long)a.at<int64_t>(i)); }
        else if (a.type == VT::F32) snprintf(buf, sizeof(buf), "%.9g", a.at<double>(i));
        else snprintf(buf, sizeof(buf), "%.17g", a.at<double>(i));
        line += (i % 6 == 0) ? std::string(indent) + "  " : " ";
        line += buf;
        if (i % 6 == 5 || i + 1 == n) { line += "\n"; xml += line; line.clear(); }
      }
      xml += std::string(indent) + "</DataArray>\n";
      return;
    }
    std::vector<unsigned char> hd, pl;
    write_block(a.bytes.data(), a.bytes.size(), hsize, compress, hd, pl);
    if (mode == B200IO_BINARY) {
      xml += ">\n" + std::string(indent) + "  ";
      b64_encode(hd.data(), hd.size(), xml);
      b64_encode(pl.data(), pl.size(), xml);
      xml += "\n" + std::string(indent) + "</DataArray>\n";
    } else if (mode == B200IO_APPENDED_RAW) {
      snprintf(buf, sizeof(buf), " offset=\"%zu\"/>\n", app_raw.size());
      xml += buf;
      app_raw.insert(app_raw.end(), hd.begin(), hd.end());
      app_raw.insert(app_raw.end(), pl.begin(), pl.end());
    } else {
      snprintf(buf, sizeof(buf), " offset=\"%zu\"/>\n", app_b64.size());
      xml += buf;
      b64_encode(hd.data(), hd.size(), app_b64);
      b64_encode(pl.data(), pl.size(), app_b64);
    }
  }
};

void write_file(const b200io_vtk& h, const char* path, int mode, int compress, int header64)
{
  if (mode < 0 || mode > 3) throw std::runtime_error("unknown data mode");
  Writer w;
  w.mode = mode; w.hsize = header64 ? 8 : 4; w.compress = compress && mode != B200IO_ASCII;
  const char* type = h.polydata ? "PolyData" : "UnstructuredGrid";
  char buf[512];
  w.xml = "<?xml version=\"1.0\"?>\n";
  snprintf(buf, sizeof(buf), "<VTKFile type=\"%s\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"%s\"%s>\n", type, header64 ? "UInt64" : "UInt32",
           w.compress ? " compressor=\"vtkZLibDataCompressor\"" : "");
  w.xml += buf;
  w.xml += std::string("  <") + type + ">\n";
  if (h.polydata)
    snprintf(buf, sizeof(buf), "    <Piece NumberOfPoints=\"%d\" NumberOfVerts=\"0\" NumberOfLines=\"0\" NumberOfStrips=\"0\" NumberOfPolys=\"%d\">\n", h.nNo, h.nEl);
  else
    snprintf(buf, sizeof(buf), "    <Piece NumberOfPoints=\"%d\" NumberOfCells=\"%d\">\n", h.nNo, h.nEl);
  w.xml += buf;
  for (int where = 0; where < 2; where++) {
    const auto& arrs = h.arrays(where);
    const char* sec = where ? "CellData" : "PointData";
    w.xml += std::string("      <") + sec + ">\n";
    for (auto& a : arrs) w.array(a, "        ");
    w.xml += std::string("      </") + sec + ">\n";
  }
  w.xml += "      <Points>\n";
  w.array(make_array("Points", VT::F64, 3, h.x.data(), h.x.size()), "        ");
  w.xml += "      </Points>\n";
  const char* csec = h.polydata ? "Polys" : "Cells";
  w.xml += std::string("      <") + csec + ">\n";
  w.array(make_array("connectivity", VT::I64, 1, h.conn.data(), h.conn.size()), "        ");
  w.array(make_array("offsets", VT::I64, 1, h.offsets.data(), h.offsets.size()), "        ");
  if (!h.polydata) w.array(make_array("types", VT::U8, 1, h.types.data(), h.types.size()), "        ");
  w.xml += std::string("      </") + csec + ">\n";
  w.xml += "    </Piece>\n";
  w.xml += std::string("  </") + type + ">\n";
  std::ofstream out(path, std::ios::binary | std::ios::trunc);
  if (!out) throw std::runtime_error(std::string("cannot open '") + path + "' for writing");
  out.write(w.xml.data(), std::streamsize(w.xml.size()));
  if (mode == B200IO_APPENDED_RAW) {
    out << "  <AppendedData encoding=\"raw\">\n   _";
    out.write((const char*)w.app_raw.data(), std::streamsize(w.app_raw.size()));
    out << "\n  </AppendedData>\n";
  } else if (mode == B200IO_APPENDED_BASE64) {
    out << "  <AppendedData encoding=\"base64\">\n   _" << w.app_b64 << "\n  </AppendedData>\n";
  }
  out << "</VTKFile>\n";
  if (!out) throw std::runtime_error(std::string("write to '") + path + "' failed");
}

const DataArray* find_array(const b200io_vtk* h, int where, const char* name)
{
  for (auto& a : h->arrays(where)) if (a.name == name) return &a;
  return nullptr;
}

template <class F> int guard(F&& f)
{
  try { f(); return 0; }
  catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// %4.3e of the reference's sprintf calls
std::string e3(double v) { char b[64]; snprintf(b, sizeof(b), "%4.3e", v); return b; }

} // namespace

extern "C" {

const char* b200io_last_error(void) { return g_err.c_str(); }

int b200io_vtk_read(const char* path, b200io_vtk** out)
{
  return guard([&] { if (!path || !out) throw std::runtime_error("null argument"); *out = read_file(path); });
}
int b200io_vtk_is_polydata(const b200io_vtk* h) { return h && h->polydata; }
int b200io_vtk_num_points(const b200io_vtk* h) { return h ? h->nNo : 0; }
int b200io_vtk_num_cells(const b200io_vtk* h) { return h ? h->nEl : 0; }
int b200io_vtk_nodes_per_cell(const b200io_vtk* h)
{
  if (!h || h->nEl == 0) return -1;
  const int64_t n = h->offsets[0];
  for (int e = 1; e < h->nEl; e++) if (h->offsets[e] - h->offsets[e-1] != n) return -1;
  return int(n);
}
int b200io_vtk_points(const b200io_vtk* h, double* x)
{
  return guard([&] { if (!h || !x) throw std::runtime_error("null argument"); std::copy(h->x.begin(), h->x.end(), x); });
}
int b200io_vtk_connectivity(const b200io_vtk* h, int* ien)
{
  return guard([&] { if (!h || !ien) throw std::runtime_error("null argument"); for (size_t i = 0; i < h->conn.size(); i++) ien[i] = int(h->conn[i]); });
}
int b200io_vtk_cell_types(const b200io_vtk* h, unsigned char* types)
{
  return guard([&] { if (!h || !types) throw std::runtime_error("null argument"); std::copy(h->types.begin(), h->types.end(), types); });
}
int b200io_vtk_num_arrays(const b200io_vtk* h, int where) { return h ? int(h->arrays(where).size()) : 0; }
const char* b200io_vtk_array_name(const b200io_vtk* h, int where, int i)
{
  if (!h || i < 0 || i >= int(h->arrays(where).size())) return nullptr;
  return h->arrays(where)[i].name.c_str();
}
int b200io_vtk_array_info(const b200io_vtk* h, int where, const char* name, int* ncomp, int* ntuples, int* is_integer)
{
  const DataArray* a = (h && name) ? find_array(h, where, name) : nullptr;
  if (!a) return 1;
  if (ncomp) *ncomp = a->ncomp;
  if (ntuples) *ntuples = int(a->count()/a->ncomp);
  if (is_integer) *is_integer = type_info(a->type).integer;
  return 0;
}
int b200io_vtk_array_f64(const b200io_vtk* h, int where, const char* name, double* out)
{
  return guard([&] {
    const DataArray* a = (h && name && out) ? find_array(h, where, name) : nullptr;
    if (!a) throw std::runtime_error(std::string("no data array named '") + (name ? name : "") + "'");
    for (size_t i = 0; i < a->count(); i++) out[i] = a->at<double>(i);
  });
}
int b200io_vtk_array_i32(const b200io_vtk* h, int where, const char* name, int* out)
{
  return guard([&] {
    const DataArray* a = (h && name && out) ? find_array(h, where, name) : nullptr;
    if (!a) throw std::runtime_error(std::string("no data array named '") + (name ? name : "") + "'");
    for (size_t i = 0; i < a->count(); i++) {
      const int64_t v = type_info(a->type).integer ? a->at<int64_t>(i) : int64_t(std::llround(a->at<double>(i)));
      if (v < std::numeric_limits<int>::min() || v > std::numeric_limits<int>::max()) throw std::runtime_error(std::string("data array '") + name + "' does not fit 32-bit integers");
      out[i] = int(v);
    }
  });
}
void b200io_vtk_free(b200io_vtk* h) { delete h; }

b200io_vtk* b200io_vtk_new(int is_polydata)
{
  auto h = new b200io_vtk;
  h->polydata = is_polydata != 0;
  return h;
}
int b200io_vtk_set_points(b200io_vtk* h, int nNo, const double* x)
{
  return guard([&] { if (!h || nNo < 0 || (nNo && !x)) throw std::runtime_error("bad argument"); h->nNo = nNo; h->x.assign(x, x + size_t(3)*nNo); });
}
int b200io_vtk_set_cells(b200io_vtk* h, int nEl, int eNoN, const int* ien, int vtk_type)
{
  return guard([&] {
    if (!h || nEl < 0 || eNoN < 1 || (nEl && !ien)) throw std::runtime_error("bad argument");
    for (size_t i = 0; i < size_t(nEl)*eNoN; i++) if (ien[i] < 0 || ien[i] >= h->nNo) throw std::runtime_error("connectivity refers to a point that does not exist (set the points first)");
    h->nEl = nEl;
    h->conn.assign(ien, ien + size_t(nEl)*eNoN);
    h->offsets.resize(nEl);
    for (int e = 0; e < nEl; e++) h->offsets[e] = int64_t(e + 1)*eNoN;
    h->types.assign(nEl, (unsigned char)vtk_type);
  });
}
int b200io_vtk_add_array_f64(b200io_vtk* h, int where, const char* name, int ncomp, int ntuples, const double* data)
{
  return guard([&] {
    if (!h || !name || ncomp < 1 || ntuples < 0 || (ntuples && !data)) throw std::runtime_error("bad argument");
    if (ntuples != (where == B200IO_CELL_DATA ? h->nEl : h->nNo)) throw std::runtime_error(std::string("data array '") + name + "': tuple count differs from the number of points / cells");
    h->arrays(where).push_back(make_array(name, VT::F64, ncomp, data, size_t(ncomp)*ntuples));
  });
}
int b200io_vtk_add_array_i32(b200io_vtk* h, int where, const char* name, int ncomp, int ntuples, const int* data)
{
  return guard([&] {
    if (!h || !name || ncomp < 1 || ntuples < 0 || (ntuples && !data)) throw std::runtime_error("bad argument");
    if (ntuples != (where == B200IO_CELL_DATA ? h->nEl : h->nNo)) throw std::runtime_error(std::string("data array '") + name + "': tuple count differs from the number of points / cells");
    h->arrays(where).push_back(make_array(name, VT::I32, ncomp, data, size_t(ncomp)*ntuples));
  });
}
int b200io_vtk_write(const b200io_vtk* h, const char* path, int mode, int compress, int header64)
{
  return guard([&] { if (!h || !path) throw std::runtime_error("null argument"); write_file(*h, path, mode, compress, header64); });
}

// ---- restart ------------------------------------------------------------------------------------------------------------
long long b200io_restart_record_bytes(const b200io_restart* r)
{
  // initialize.cpp:505-513
  long long i = 2LL*r->tDof;
  if (r->dFlag) i = 3LL*r->tDof;
  if (r->pstEq) i += r->nsymd;
  if (r->sstEq) i += r->nsd;
  return (long long)sizeof(int)*(1 + 7) + (long long)sizeof(double)*(2 + r->nEq + r->nXn + i*r->tnNo);
}

int b200io_restart_name(const char* stem, int cTS, char* out, int cap)
{
  char num[32];
  if (cTS >= 1000) snprintf(num, sizeof(num), "%d", cTS); else snprintf(num, sizeof(num), "%03d", cTS);
  return snprintf(out, size_t(cap), "%s_%s.bin", stem, num);
}

int b200io_restart_write(const char* path, int rank, long long recLn, const b200io_restart* r, int create)
{
  return guard([&] {
    if (!path || !r || rank < 0) throw std::runtime_error("bad argument");
    if (create) { std::ofstream c(path, std::ios::out | std::ios::binary); if (!c) throw std::runtime_error(std::string("cannot create '") + path + "'"); }
    std::ofstream f(path, std::ios::out | std::ios::binary | std::ios::in);
    if (!f) throw std::runtime_error(std::string("cannot open '") + path + "' for writing");
    f.seekp(std::streamoff(rank)*recLn);
    const size_t nv = size_t(r->tDof)*r->tnNo*sizeof(double);
    f.write((const char*)r->stamp, sizeof(r->stamp));
    f.write((const char*)&r->cTS, sizeof(int));
    f.write((const char*)&r->time, sizeof(double));
    f.write((const char*)&r->cpu_time, sizeof(double));
    f.write((const char*)r->iNorm, std::streamsize(sizeof(double)*r->nEq));
    f.write((const char*)r->xn, std::streamsize(sizeof(double)*r->nXn));
    f.write((const char*)r->Yn, std::streamsize(nv));
    f.write((const char*)r->An, std::streamsize(nv));
    if (r->dFlag) {
      f.write((const char*)r->Dn, std::streamsize(nv));
      if (r->sstEq) {
        if (r->pstEq) f.write((const char*)r->pS0, std::streamsize(sizeof(double)*r->nsymd*r->tnNo));
        f.write((const char*)r->Ad, std::streamsize(sizeof(double)*r->nsd*r->tnNo));
      } else if (r->pstEq) f.write((const char*)r->pS0, std::streamsize(sizeof(double)*r->nsymd*r->tnNo));
      else if (r->trailing_Dn) f.write((const char*)r->Dn, std::streamsize(nv));
    }
    if (!f) throw std::runtime_error(std::string("write to '") + path + "' failed");
  });
}

int b200io_restart_read(const char* path, int rank, long long recLn, b200io_restart* r)
{
  return guard([&] {
    if (!path || !r || rank < 0) throw std::runtime_error("bad argument");
    std::ifstream f(path, std::ios::binary | std::ios::in);
    if (!f) throw std::runtime_error(std::string("cannot open '") + path + "'");
    f.seekg(std::streamoff(rank)*recLn);
    const size_t nv = size_t(r->tDof)*r->tnNo*sizeof(double);
    f.read((char*)r->stamp, sizeof(r->stamp));
    f.read((char*)&r->cTS, sizeof(int));
    f.read((char*)&r->time, sizeof(double));
    f.read((char*)&r->cpu_time, sizeof(double));
    f.read((char*)r->iNorm, std::streamsize(sizeof(double)*r->nEq));
    f.read((char*)r->xn, std::streamsize(sizeof(double)*r->nXn));
    f.read((char*)r->Yn, std::streamsize(nv));
    f.read((char*)r->An, std::streamsize(nv));
    if (r->dFlag) {
      f.read((char*)r->Dn, std::streamsize(nv));
      if (r->sstEq) {
        if (r->pstEq) f.read((char*)r->pS0, std::streamsize(sizeof(double)*r->nsymd*r->tnNo));
        f.read((char*)r->Ad, std::streamsize(sizeof(double)*r->nsd*r->tnNo));
      } else if (r->pstEq) f.read((char*)r->pS0, std::streamsize(sizeof(double)*r->nsymd*r->tnNo));
    }
    if (!f) throw std::runtime_error(std::string("'") + path + "' is shorter than the record of rank " + std::to_string(rank));
  });
}

// ---- history ------------------------------------------------------------------------------------------------------------
int b200io_history_header(int nEq, char* out, int cap)
{
  const std::string sep(69, '-');
  std::string s = sep + "\n" + " Eq     N-i     T       dB  Ri/R1   Ri/R0    R/Ri     lsIt   dB  %t" + "\n";
  if (nEq == 1) s += sep + "\n";
  return snprintf(out, size_t(cap), "%s", s.c_str());
}

int b200io_history_line(const b200io_history* h, char* out, int cap)
{
  // output.cpp:88-165
  std::string c1 = h->saved ? "s" : " ";
  std::string s = std::string(" ") + h->sym + " " + std::to_string(h->cTS) + "-" + std::to_string(h->itr) + c1 + " " + e3(h->elapsed);
  double tmp, tmp1, tmp2; int i;
  const double eps = std::numeric_limits<double>::epsilon();
  // utils::is_zero(eq.iNorm) (utils.cpp:170-188): |v| / max(|v|, eps) < 10 eps
  if (std::fabs(h->eq_iNorm)/std::fmax(std::fabs(h->eq_iNorm), eps) < 10.0*eps) { tmp = tmp1 = tmp2 = 1.0; i = 0; }
  else {
    tmp = h->ri_iNorm/h->eq_iNorm;
    tmp1 = tmp/h->eq_pNorm;
    tmp2 = h->ri_fNorm/h->ri_iNorm;
    i = int(20.0*log10(tmp1));
  }
  std::string c2;
  if (i > 20) { c1 = "!"; c2 = "!"; } else { c1 = "["; c2 = "]"; }
  s += "  " + c1 + std::to_string(i) + " " + e3(tmp1) + " " + e3(tmp) + " " + e3(tmp2) + c2;
  // a zero interval is widened like output.cpp:134-136 does
  const double since = (h->since_last == 0.0) ? eps : h->since_last;
  double pct = 100.0*h->ri_callD/since;
  if (std::fabs(pct) > 100.0) pct = 100.0;
  std::string warn;
  if (h->ri_suc) { c1 = "["; c2 = "]"; }
  else { c1 = "!"; c2 = "!"; warn = "  WARNING: The linear system solution has not converged"; }
  s += "  " + c1 + std::to_string(h->ri_itr) + " " + std::to_string(int(std::round(h->ri_dB))) + " " + std::to_string(int(std::round(pct))) + c2 + warn;
  return snprintf(out, size_t(cap), "%s\n", s.c_str());
}

} // extern "C"
