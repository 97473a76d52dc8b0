// Dependency-free reader / writer for the reference's grid file `*.msh.h5`
// (load_grid_gmsh_h5, src/zisa/grid/grid.cpp:889-901; written by src/renumber_grid.cpp:129-132): three datasets in the
// root group -- the scalar `n_dims`, `vertex_indices` [n_cells][n_dims + 1] and `vertices` [n_vertices][3].
//
// No HDF5 library exists in this image, so this file implements the subset of the HDF5 File Format Specification
// (version 3.0) those files use, nothing more:
//   superblock 0 / 1 (root group = symbol table: v1 B-tree "TREE" -> "SNOD" nodes, names in a local heap "HEAP") and
//   superblock 2 / 3 (root object header with compact Link messages);
//   object headers version 1 and 2 ("OHDR" / "OCHK") with continuation blocks;
//   messages: dataspace (v1, v2), datatype (fixed point 1/2/4/8 bytes, IEEE float 4/8 bytes, little endian),
//   data layout v3 / v4 (contiguous or compact), symbol table, link.
// Chunked / compressed datasets, dense link storage, big-endian files and everything else are rejected with a message
// that says which feature the file uses.  The writer emits the layout libhdf5 1.8 / h5py write by default
// (superblock 0, 8-byte offsets, symbol-table root group, version-1 object headers, contiguous datasets).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "zfvm_host.hpp"

namespace zfvm {

namespace {

using u8 = std::uint8_t;
using u64 = std::uint64_t;
constexpr u64 UNDEF = ~0ull;
const u8 SIGNATURE[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};

struct Dataset {
  int rank = -1;               // 0 = scalar
  u64 dims[4] = {0, 0, 0, 0};
  int type_class = -1;         // 0 fixed point, 1 floating point
  int type_size = 0;
  bool is_signed = false;
  bool have_layout = false;
  u64 address = UNDEF, size = 0;   // contiguous
  std::vector<u8> compact;         // compact layout
  bool is_compact = false;
  u64 count() const {
    u64 c = 1;
    for (int d = 0; d < rank; ++d) c *= dims[d];
    return c;
  }
};

struct Link {
  std::string name;
  u64 header = UNDEF;
};

class File {
 public:
  explicit File(const std::string &path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open " + path);
    in.seekg(0, std::ios::end);
    const std::streamoff len = in.tellg();
    in.seekg(0);
    buf_.resize((size_t)len);
    if (len > 0) in.read(reinterpret_cast<char *>(buf_.data()), len);
    if (!in) throw std::runtime_error("cannot read " + path);
    parse_superblock();
  }

  bool find(const std::string &name, Dataset &ds) {
    for (const Link &l : links_)
      if (l.name == name) {
        ds = read_dataset(l.header);
        return true;
      }
    return false;
  }

  /// Elements of a dataset converted to T (integers of any width, float / double).
  template <class T>
  std::vector<T> values(const Dataset &ds, const std::string &name) {
    const u64 n = ds.count();
    const u8 *src;
    if (ds.is_compact) {
      if (ds.compact.size() < n * (u64)ds.type_size) throw std::runtime_error(name + ": compact data too short");
      src = ds.compact.data();
    } else {
      if (n == 0) return {};
      if (ds.address == UNDEF) throw std::runtime_error(name + ": dataset has no storage allocated");
      src = at(base_ + ds.address, n * (u64)ds.type_size);
    }
    std::vector<T> out((size_t)n);
    for (u64 i = 0; i < n; ++i) {
      const u8 *p = src + i * (u64)ds.type_size;
      if (ds.type_class == 1) {
        if (ds.type_size == 8) {
          double v;
          std::memcpy(&v, p, 8);
          out[(size_t)i] = (T)v;
        } else {
          float v;
          std::memcpy(&v, p, 4);
          out[(size_t)i] = (T)v;
        }
      } else {
        u64 raw = 0;
        std::memcpy(&raw, p, (size_t)ds.type_size);
        if (ds.is_signed && ds.type_size < 8 && (raw >> (8 * ds.type_size - 1)) & 1) raw |= ~0ull << (8 * ds.type_size);
        out[(size_t)i] = ds.is_signed ? (T)(std::int64_t)raw : (T)raw;
      }
    }
    return out;
  }

 private:
  std::vector<u8> buf_;
  u64 base_ = 0;
  int so_ = 8, sl_ = 8;  // size of offsets / lengths
  std::vector<Link> links_;

  const u8 *at(u64 off, u64 len) const {
    if (off > buf_.size() || len > buf_.size() - off) throw std::runtime_error("address beyond the end of the file");
    return buf_.data() + off;
  }
  u64 rd(u64 off, int n) const {
    u64 v = 0;
    std::memcpy(&v, at(off, (u64)n), (size_t)n);
    return v;
  }
  u64 rd_off(u64 off) const {
    const u64 v = rd(off, so_);
    return (so_ < 8 && v == (UNDEF >> (64 - 8 * so_))) ? UNDEF : v;
  }

  void parse_superblock() {
    // the superblock sits at 0, 512, 1024, ... (a user block may precede it)
    u64 sb = UNDEF;
    for (u64 off = 0; off + 8 <= buf_.size(); off = off ? 2 * off : 512)
      if (std::memcmp(buf_.data() + off, SIGNATURE, 8) == 0) {
        sb = off;
        break;
      }
    if (sb == UNDEF) throw std::runtime_error("not an HDF5 file (no signature)");
    const int version = (int)rd(sb + 8, 1);
    u64 root_header = UNDEF;
    if (version == 0 || version == 1) {
      so_ = (int)rd(sb + 13, 1);
      sl_ = (int)rd(sb + 14, 1);
      check_sizes();
      u64 p = sb + 24 + (version == 1 ? 4 : 0);
      base_ = rd_off(p);
      p += 4 * (u64)so_;          // base, free-space info, end of file, driver info
      root_header = rd_off(p + (u64)so_);  // symbol table entry: link name offset, object header address
    } else if (version == 2 || version == 3) {
      so_ = (int)rd(sb + 9, 1);
      sl_ = (int)rd(sb + 10, 1);
      check_sizes();
      base_ = rd_off(sb + 12);
      root_header = rd_off(sb + 12 + 3 * (u64)so_);
    } else {
      throw std::runtime_error("unsupported superblock version " + std::to_string(version));
    }
    if (base_ == UNDEF) base_ = 0;
    if (root_header == UNDEF) throw std::runtime_error("file has no root group");
    read_group(root_header);
  }
  void check_sizes() const {
    if ((so_ != 4 && so_ != 8) || (sl_ != 4 && sl_ != 8)) throw std::runtime_error("unsupported offset / length size");
  }

  // ---- object headers: calls f(type, data offset, data size) for every message -----------------------------------
  template <class F>
  void for_each_message(u64 header, F &&f) {
    const u64 h = base_ + header;
    if (std::memcmp(at(h, 4), "OHDR", 4) == 0) {
      if (rd(h + 4, 1) != 2) throw std::runtime_error("unsupported object header version");
      const int flags = (int)rd(h + 5, 1);
      u64 p = h + 6;
      if (flags & 0x20) p += 16;
      if (flags & 0x10) p += 4;
      const int szb = 1 << (flags & 3);
      const u64 chunk0 = rd(p, szb);
      p += (u64)szb;
      messages_v2(p, chunk0, (flags & 4) != 0, f, 0);
    } else {
      if (rd(h, 1) != 1) throw std::runtime_error("unsupported object header version");
      const u64 size = rd(h + 8, 4);
      messages_v1(h + 16, size, f, 0);
    }
  }
  template <class F>
  void messages_v1(u64 p, u64 size, F &&f, int depth) {
    if (depth > 64) throw std::runtime_error("object header continuation loop");
    const u64 end = p + size;
    while (p + 8 <= end) {
      const int type = (int)rd(p, 2);
      const u64 msize = rd(p + 2, 2);
      const u64 data = p + 8;
      if (data + msize > end) throw std::runtime_error("object header message overruns its block");
      if (type == 0x10)
        messages_v1(base_ + rd_off(data), rd(data + (u64)so_, sl_), f, depth + 1);
      else
        f(type, data, msize);
      p = data + msize;
    }
  }
  template <class F>
  void messages_v2(u64 p, u64 size, bool creation_order, F &&f, int depth) {
    if (depth > 64) throw std::runtime_error("object header continuation loop");
    const u64 end = p + size;
    const u64 hdr = 4 + (creation_order ? 2 : 0);
    while (p + hdr <= end) {
      const int type = (int)rd(p, 1);
      const u64 msize = rd(p + 1, 2);
      const u64 data = p + hdr;
      if (data + msize > end) break;  // gap before the checksum
      if (type == 0x10) {
        const u64 block = base_ + rd_off(data), blen = rd(data + (u64)so_, sl_);
        if (std::memcmp(at(block, 4), "OCHK", 4) != 0) throw std::runtime_error("bad continuation block signature");
        messages_v2(block + 4, blen - 8, creation_order, f, depth + 1);
      } else {
        f(type, data, msize);
      }
      p = data + msize;
    }
  }

  // ---- root group -----------------------------------------------------------------------------------------------------
  void read_group(u64 header) {
    u64 btree = UNDEF, heap = UNDEF;
    bool dense = false;
    for_each_message(header, [&](int type, u64 data, u64 size) {
      if (type == 0x11) {  // symbol table
        btree = rd_off(data);
        heap = rd_off(data + (u64)so_);
      } else if (type == 0x06) {  // link
        read_link(data, size);
      } else if (type == 0x02) {  // link info: dense storage when the fractal heap address is defined
        const int flags = (int)rd(data + 1, 1);
        u64 p = data + 2 + ((flags & 1) ? 8 : 0);
        if (rd_off(p) != UNDEF) dense = true;
      }
    });
    if (btree != UNDEF) {
      const u64 hp = base_ + heap;
      if (std::memcmp(at(hp, 4), "HEAP", 4) != 0) throw std::runtime_error("bad local heap signature");
      const u64 heap_data = base_ + rd_off(hp + 8 + 2 * (u64)sl_);
      walk_btree(btree, heap_data, 0);
    } else if (links_.empty() && dense) {
      throw std::runtime_error("the root group uses dense link storage (fractal heap), which this reader does not implement");
    }
  }
  void read_link(u64 p, u64 size) {
    const u64 end = p + size;
    if (rd(p, 1) != 1) throw std::runtime_error("unsupported link message version");
    const int flags = (int)rd(p + 1, 1);
    p += 2;
    int type = 0;
    if (flags & 8) type = (int)rd(p++, 1);
    if (flags & 4) p += 8;
    if (flags & 16) p += 1;
    const int nb = 1 << (flags & 3);
    const u64 nlen = rd(p, nb);
    p += (u64)nb;
    if (p + nlen > end) throw std::runtime_error("link message overruns");
    Link l;
    l.name.assign(reinterpret_cast<const char *>(at(p, nlen)), (size_t)nlen);
    p += nlen;
    if (type != 0) return;  // soft / external links are not followed
    l.header = rd_off(p);
    links_.push_back(l);
  }
  void walk_btree(u64 node, u64 heap_data, int depth) {
    if (depth > 32) throw std::runtime_error("group B-tree too deep");
    const u64 p = base_ + node;
    if (std::memcmp(at(p, 4), "TREE", 4) != 0) throw std::runtime_error("bad B-tree node signature");
    if (rd(p + 4, 1) != 0) throw std::runtime_error("not a group B-tree");
    const int level = (int)rd(p + 5, 1);
    const int used = (int)rd(p + 6, 2);
    u64 q = p + 8 + 2 * (u64)so_;  // keys and children alternate: key 0, child 0, key 1, ...
    for (int i = 0; i < used; ++i) {
      const u64 child = rd_off(q + (u64)sl_);
      q += (u64)sl_ + (u64)so_;
      if (level > 0) {
        walk_btree(child, heap_data, depth + 1);
        continue;
      }
      const u64 s = base_ + child;
      if (std::memcmp(at(s, 4), "SNOD", 4) != 0) throw std::runtime_error("bad symbol table node signature");
      const int n_sym = (int)rd(s + 6, 2);
      const u64 entry = 2 * (u64)so_ + 24;
      for (int e = 0; e < n_sym; ++e) {
        const u64 ep = s + 8 + (u64)e * entry;
        Link l;
        const u64 name_off = rd_off(ep);
        const char *nm = reinterpret_cast<const char *>(at(heap_data + name_off, 1));
        const u64 max_len = buf_.size() - (heap_data + name_off);
        l.name.assign(nm, strnlen(nm, (size_t)max_len));
        l.header = rd_off(ep + (u64)so_);
        links_.push_back(l);
      }
    }
  }

  // ---- datasets -------------------------------------------------------------------------------------------------------
  Dataset read_dataset(u64 header) {
    Dataset ds;
    for_each_message(header, [&](int type, u64 data, u64 size) {
      if (type == 0x01) {  // dataspace
        const int version = (int)rd(data, 1);
        ds.rank = (int)rd(data + 1, 1);
        if (ds.rank > 4) throw std::runtime_error("dataset rank above 4");
        u64 p;
        if (version == 1) {
          p = data + 8;
        } else if (version == 2) {
          if (rd(data + 3, 1) == 2) throw std::runtime_error("null dataspace");
          p = data + 4;
        } else {
          throw std::runtime_error("unsupported dataspace version");
        }
        for (int d = 0; d < ds.rank; ++d) ds.dims[d] = rd(p + (u64)d * (u64)sl_, sl_);
      } else if (type == 0x03) {  // datatype
        const int cv = (int)rd(data, 1);
        ds.type_class = cv & 15;
        const int bits0 = (int)rd(data + 1, 1);
        ds.type_size = (int)rd(data + 4, 4);
        if (ds.type_class == 0) {
          ds.is_signed = (bits0 & 8) != 0;
          if (bits0 & 1) throw std::runtime_error("big-endian integers are not supported");
          if (ds.type_size != 1 && ds.type_size != 2 && ds.type_size != 4 && ds.type_size != 8)
            throw std::runtime_error("unsupported integer size");
        } else if (ds.type_class == 1) {
          if (bits0 & 1) throw std::runtime_error("big-endian floats are not supported");
          if (ds.type_size != 4 && ds.type_size != 8) throw std::runtime_error("unsupported float size");
        } else {
          throw std::runtime_error("unsupported datatype class " + std::to_string(ds.type_class) +
                                   " (only integers and IEEE floats)");
        }
      } else if (type == 0x08) {  // data layout
        const int version = (int)rd(data, 1);
        if (version != 3 && version != 4) throw std::runtime_error("unsupported data layout message version");
        const int cls = (int)rd(data + 1, 1);
        if (cls == 1) {
          ds.address = rd_off(data + 2);
          ds.size = rd(data + 2 + (u64)so_, sl_);
        } else if (cls == 0) {
          const u64 n = rd(data + 2, 2);
          const u8 *p = at(data + 4, n);
          ds.compact.assign(p, p + n);
          ds.is_compact = true;
        } else {
          throw std::runtime_error("chunked / virtual dataset layouts are not supported (rewrite the file contiguous, "
                                   "without compression)");
        }
        ds.have_layout = true;
      } else if (type == 0x0B) {
        if (size > 0 && rd(data, 1) >= 1) {
          const int nfilters = (int)rd(data + 1, 1);
          if (nfilters > 0) throw std::runtime_error("filtered (compressed) datasets are not supported");
        }
      }
    });
    if (ds.rank < 0 || ds.type_class < 0 || !ds.have_layout) throw std::runtime_error("object is not a simple dataset");
    return ds;
  }
};

// ---- writer ---------------------------------------------------------------------------------------------------------------
struct Out {
  std::vector<u8> b;
  u64 size() const { return b.size(); }
  void put(u64 v, int n) {
    for (int i = 0; i < n; ++i) b.push_back((u8)(v >> (8 * i)));
  }
  void bytes(const void *p, size_t n) {
    const u8 *q = static_cast<const u8 *>(p);
    b.insert(b.end(), q, q + n);
  }
  void pad8() {
    while (b.size() % 8) b.push_back(0);
  }
  void patch(u64 at_, u64 v, int n) {
    for (int i = 0; i < n; ++i) b[(size_t)at_ + i] = (u8)(v >> (8 * i));
  }
};

struct WDataset {
  std::string name;
  int rank;
  u64 dims[2];
  int type_class, type_size;
  bool is_signed;
  const void *data;
  u64 bytes;
};

void message_v1(Out &o, int type, const Out &body) {
  o.put((u64)type, 2);
  const u64 padded = (body.size() + 7) / 8 * 8;
  o.put(padded, 2);
  o.put(0, 1);
  o.put(0, 3);
  o.bytes(body.b.data(), body.b.size());
  for (u64 i = body.size(); i < padded; ++i) o.put(0, 1);
}

/// Version-1 object header of a contiguous dataset whose raw data sits at `data_address`.
void dataset_header(Out &o, const WDataset &d, u64 data_address) {
  Out msgs;
  {
    Out m;  // dataspace, version 1
    m.put(1, 1);
    m.put((u64)d.rank, 1);
    m.put(0, 1);
    m.put(0, 1);
    m.put(0, 4);
    for (int k = 0; k < d.rank; ++k) m.put(d.dims[k], 8);
    message_v1(msgs, 0x01, m);
  }
  {
    Out m;  // datatype, version 1
    m.put((u64)(0x10 | d.type_class), 1);
    if (d.type_class == 0) {
      m.put(d.is_signed ? 0x08 : 0x00, 1);
      m.put(0, 2);
      m.put((u64)d.type_size, 4);
      m.put(0, 2);
      m.put((u64)(8 * d.type_size), 2);
    } else {  // IEEE double, little endian
      m.put(0x20, 1);
      m.put(0x3f, 1);
      m.put(0, 1);
      m.put(8, 4);
      m.put(0, 2);
      m.put(64, 2);
      m.put(52, 1);
      m.put(11, 1);
      m.put(0, 1);
      m.put(52, 1);
      m.put(1023, 4);
    }
    message_v1(msgs, 0x03, m);
  }
  {
    Out m;  // fill value, version 2: late allocation, written if set, default value
    m.put(2, 1);
    m.put(2, 1);
    m.put(2, 1);
    m.put(1, 1);
    m.put(0, 4);
    message_v1(msgs, 0x05, m);
  }
  {
    Out m;  // data layout, version 3, contiguous
    m.put(3, 1);
    m.put(1, 1);
    m.put(d.bytes ? data_address : UNDEF, 8);
    m.put(d.bytes, 8);
    message_v1(msgs, 0x08, m);
  }
  o.put(1, 1);
  o.put(0, 1);
  o.put(4, 2);
  o.put(1, 4);
  o.put(msgs.size(), 4);
  o.put(0, 4);  // messages start on an 8-byte boundary
  o.bytes(msgs.b.data(), msgs.b.size());
}

}  // namespace

void read_msh_h5(const std::string &path, RawMesh &mesh) {
  File f(path);
  Dataset nd, vi, vx;
  if (!f.find("n_dims", nd)) throw std::runtime_error(path + ": no dataset 'n_dims'");
  if (!f.find("vertex_indices", vi)) throw std::runtime_error(path + ": no dataset 'vertex_indices'");
  if (!f.find("vertices", vx)) throw std::runtime_error(path + ": no dataset 'vertices'");
  const std::vector<std::int64_t> nd_v = f.values<std::int64_t>(nd, "n_dims");
  if (nd_v.size() != 1 || (nd_v[0] != 2 && nd_v[0] != 3)) throw std::runtime_error(path + ": n_dims must be 2 or 3");
  mesh.n_dims = (int)nd_v[0];
  const int F = mesh.n_dims + 1;
  if (vi.type_class != 0 || vi.rank != 2 || vi.dims[1] != (u64)F)
    throw std::runtime_error(path + ": vertex_indices must be an integer array [n_cells][n_dims + 1]");
  if (vx.type_class != 1 || !((vx.rank == 2 && vx.dims[1] == 3) || (vx.rank == 1 && vx.dims[0] % 3 == 0)))
    throw std::runtime_error(path + ": vertices must be a floating-point array [n_vertices][3]");
  mesh.vertices = f.values<double>(vx, "vertices");
  const std::vector<std::int64_t> idx = f.values<std::int64_t>(vi, "vertex_indices");
  const std::int64_t n_vertices = (std::int64_t)mesh.vertices.size() / 3;
  mesh.vertex_indices.resize(idx.size());
  for (size_t a = 0; a < idx.size(); ++a) {
    if (idx[a] < 0 || idx[a] >= n_vertices) throw std::runtime_error(path + ": vertex index out of range");
    mesh.vertex_indices[a] = (i32)idx[a];
  }
  // sub-grid files (subgrid-%04d.msh.h5, src/domain_decomposition.cpp:84-88) carry two more datasets of int_t per cell
  Dataset pt, gc;
  const bool has_pt = f.find("partition", pt), has_gc = f.find("global_cell_indices", gc);
  mesh.partition.clear();
  mesh.global_cell_indices.clear();
  if (has_pt != has_gc) throw std::runtime_error(path + ": 'partition' and 'global_cell_indices' come together (sub-grid files)");
  if (has_pt) {
    const u64 n_cells = vi.dims[0];
    if (pt.type_class != 0 || pt.rank != 1 || pt.dims[0] != n_cells || gc.type_class != 0 || gc.rank != 1 || gc.dims[0] != n_cells)
      throw std::runtime_error(path + ": partition / global_cell_indices must be integer arrays [n_cells]");
    mesh.partition = f.values<std::int64_t>(pt, "partition");
    mesh.global_cell_indices = f.values<std::int64_t>(gc, "global_cell_indices");
    for (size_t a = 0; a < mesh.partition.size(); ++a)
      if (mesh.partition[a] < 0 || mesh.global_cell_indices[a] < 0)
        throw std::runtime_error(path + ": negative partition / global cell index");
  }
}

void write_msh_h5(const std::string &path, const RawMesh &mesh) {
  const int F = mesh.n_dims + 1;
  const std::int32_t n_dims = mesh.n_dims;
  // the reference's int_t is std::size_t: 64-bit unsigned indices
  std::vector<u64> idx(mesh.vertex_indices.begin(), mesh.vertex_indices.end());
  const bool sub = !mesh.partition.empty();
  if (sub && (mesh.partition.size() != idx.size() / (size_t)F || mesh.global_cell_indices.size() != mesh.partition.size()))
    throw std::runtime_error("write_msh_h5: partition / global_cell_indices must have one entry per cell");
  std::vector<u64> part(mesh.partition.begin(), mesh.partition.end()), gci(mesh.global_cell_indices.begin(), mesh.global_cell_indices.end());
  std::vector<WDataset> sets;  // in strcmp order, as a symbol table node wants them
  if (sub) sets.push_back({"global_cell_indices", 1, {(u64)gci.size(), 0}, 0, 8, false, gci.data(), gci.size() * 8});
  sets.push_back({"n_dims", 0, {0, 0}, 0, 4, true, &n_dims, 4});
  if (sub) sets.push_back({"partition", 1, {(u64)part.size(), 0}, 0, 8, false, part.data(), part.size() * 8});
  sets.push_back({"vertex_indices", 2, {(u64)(idx.size() / (size_t)F), (u64)F}, 0, 8, false, idx.data(), idx.size() * 8});
  sets.push_back({"vertices", 2, {(u64)(mesh.vertices.size() / 3), 3}, 1, 8, false, mesh.vertices.data(), mesh.vertices.size() * 8});
  const int NSETS = (int)sets.size();  // <= 2 K = 8 entries of one symbol table node

  Out o;
  // superblock, version 0 (96 bytes)
  o.bytes(SIGNATURE, 8);
  o.put(0, 1);
  o.put(0, 1);
  o.put(0, 1);
  o.put(0, 1);
  o.put(0, 1);
  o.put(8, 1);
  o.put(8, 1);
  o.put(0, 1);
  o.put(4, 2);   // group leaf node K
  o.put(16, 2);  // group internal node K
  o.put(0, 4);
  o.put(0, 8);      // base address
  o.put(UNDEF, 8);  // free-space info
  const u64 at_eof = o.size();
  o.put(0, 8);      // end of file, patched below
  o.put(UNDEF, 8);  // driver info
  o.put(0, 8);      // root entry: link name offset
  const u64 at_root_header = o.size();
  o.put(0, 8);
  o.put(1, 4);  // cache type 1: B-tree and heap addresses in the scratch pad
  o.put(0, 4);
  const u64 at_root_scratch = o.size();
  o.put(0, 8);
  o.put(0, 8);

  // root group object header: one symbol table message
  const u64 root_header = o.size();
  o.put(1, 1);
  o.put(0, 1);
  o.put(1, 2);
  o.put(1, 4);
  o.put(24, 4);
  o.put(0, 4);
  o.put(0x11, 2);
  o.put(16, 2);
  o.put(0, 1);
  o.put(0, 3);
  const u64 at_stab = o.size();
  o.put(0, 8);
  o.put(0, 8);

  // local heap: "" at offset 0, then the names, 8-byte aligned
  u64 name_off[8];
  Out heap_data;
  heap_data.put(0, 8);
  for (int k = 0; k < NSETS; ++k) {
    name_off[k] = heap_data.size();
    heap_data.bytes(sets[k].name.c_str(), sets[k].name.size() + 1);
    heap_data.pad8();
  }
  const u64 free_off = heap_data.size();
  heap_data.put(1, 8);   // free block: no next block (H5HL_FREE_NULL)
  heap_data.put(32, 8);  // its size, this header included
  heap_data.put(0, 8);
  heap_data.put(0, 8);
  const u64 heap = o.size();
  o.bytes("HEAP", 4);
  o.put(0, 1);
  o.put(0, 3);
  o.put(heap_data.size(), 8);
  o.put(free_off, 8);
  o.put(heap + 32, 8);
  o.bytes(heap_data.b.data(), heap_data.b.size());

  // B-tree node (level 0, one child) and its symbol table node
  const u64 btree = o.size();
  const u64 btree_bytes = 24 + 33 * 8 + 32 * 8, snod_bytes = 8 + 8 * 40;
  const u64 snod = btree + btree_bytes;
  o.bytes("TREE", 4);
  o.put(0, 1);
  o.put(0, 1);
  o.put(1, 2);
  o.put(UNDEF, 8);
  o.put(UNDEF, 8);
  o.put(0, 8);             // key 0: ""
  o.put(snod, 8);          // child 0
  o.put(name_off[NSETS - 1], 8);   // key 1: the largest name in child 0
  while (o.size() < btree + btree_bytes) o.put(0, 1);
  o.bytes("SNOD", 4);
  o.put(1, 1);
  o.put(0, 1);
  o.put((u64)NSETS, 2);
  u64 at_entry_header[8];
  for (int k = 0; k < NSETS; ++k) {
    o.put(name_off[k], 8);
    at_entry_header[k] = o.size();
    o.put(0, 8);
    o.put(0, 4);
    o.put(0, 4);
    o.put(0, 16);
  }
  while (o.size() < snod + snod_bytes) o.put(0, 1);

  // dataset headers, then the raw data
  u64 header[8], at_layout_addr[8];
  for (int k = 0; k < NSETS; ++k) {
    o.pad8();
    header[k] = o.size();
    dataset_header(o, sets[k], 0);
    at_layout_addr[k] = o.size() - 22;  // the layout message comes last: version, class, address, size, 6 bytes of padding
  }
  for (int k = 0; k < NSETS; ++k) {
    o.pad8();
    const u64 addr = o.size();
    o.bytes(sets[k].data, (size_t)sets[k].bytes);
    if (sets[k].bytes) o.patch(at_layout_addr[k], addr, 8);
    o.patch(at_entry_header[k], header[k], 8);
  }
  o.patch(at_eof, o.size(), 8);
  o.patch(at_root_header, root_header, 8);
  o.patch(at_root_scratch, btree, 8);
  o.patch(at_root_scratch + 8, heap, 8);
  o.patch(at_stab, btree, 8);
  o.patch(at_stab + 8, heap, 8);

  std::ofstream out(path, std::ios::binary | std::ios::trunc);
  if (!out) throw std::runtime_error("cannot create " + path);
  out.write(reinterpret_cast<const char *>(o.b.data()), (std::streamsize)o.b.size());
  if (!out) throw std::runtime_error("cannot write " + path);
}

}  // namespace zfvm
