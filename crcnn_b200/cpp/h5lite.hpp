// Minimal reader for the HDF5 files the reference's weights travel in (PlainModel/*.h5, written by h5py from a
// PyTorch state dict: PlainModel/ToH5.py).  Replaces, for this path only, the LoadH5 class of CrCNN/src/H5Easy.h:67-140
// (libhdf5 C++ API, CrCNN/src/H5Easy.cpp:584-644) behind the same member names -- setFileName / setVarName / getData /
// getSize -- so CnnBuilder::getPretrained (CrCNN/src/cnnBuilder.cpp:20-23) reads the same floats without libhdf5.
//
// Understood subset of the format (HDF5 File Format Specification, version 0 superblock, which is what h5py's default
// libver='earliest' writes): superblock v0 -> root symbol-table entry -> group B-tree v1 ("TREE") -> symbol nodes ("SNOD")
// with names in a local heap ("HEAP") -> version-1 object headers with continuation blocks -> dataspace (versions 1, 2),
// datatype (IEEE float 32/64, integers 8..64 bit, either byte order) and data-layout version 3 (contiguous or compact)
// messages.  Anything else (chunked / filtered datasets, nested groups reached by path, new-style groups) is reported
// with std::runtime_error rather than guessed at.
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iterator>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace crcnn_b200 {

struct H5Dataset {
    std::vector<std::uint64_t> dims;
    int type_class = -1;   // 0 integer, 1 float
    int type_size = 0;     // bytes per element
    bool big_endian = false, is_signed = true;
    std::uint64_t address = 0, bytes = 0;  // contiguous payload
    std::vector<unsigned char> compact;    // compact payload
    std::uint64_t count() const {
        std::uint64_t c = 1;
        for (auto d : dims) c *= d;
        return c;
    }
};

class H5File {
public:
    explicit H5File(const std::string &path) {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw std::runtime_error("h5lite: cannot open " + path);
        buf_.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
        parse();
    }
    const std::map<std::string, H5Dataset> &datasets() const { return ds_; }
    bool has(const std::string &name) const { return ds_.count(name) != 0; }

    // element values converted to double (exact for float32 and for integers below 2^53)
    std::vector<double> read(const std::string &name) const {
        auto it = ds_.find(name);
        if (it == ds_.end()) throw std::runtime_error("h5lite: no dataset named " + name);
        const H5Dataset &d = it->second;
        // sizes come from the file: nothing below trusts them (element size in {1,2,4,8}, element count x size computed without wrap-around
        // and compared with what is really there)
        if (d.type_size != 1 && d.type_size != 2 && d.type_size != 4 && d.type_size != 8)
            throw std::runtime_error("h5lite: unsupported element size " + std::to_string(d.type_size) + " in " + name);
        if (d.type_class == 1 && d.type_size < 4) throw std::runtime_error("h5lite: unsupported float size in " + name);
        std::uint64_t cnt = 1;
        for (auto dim : d.dims) {
            if (dim != 0 && cnt > (std::uint64_t)buf_.size() / dim) throw std::runtime_error("h5lite: dataset dimensions exceed the file: " + name);
            cnt *= dim;
        }
        if (cnt > (std::uint64_t)buf_.size() / (unsigned)d.type_size) throw std::runtime_error("h5lite: dataset payload beyond end of file: " + name);
        const std::uint64_t need_bytes = cnt * (unsigned)d.type_size;
        const unsigned char *p;
        if (!d.compact.empty()) {
            if (need_bytes > d.compact.size()) throw std::runtime_error("h5lite: compact payload shorter than the dataspace: " + name);
            p = d.compact.data();
        } else {
            if (cnt == 0) return {};
            if (d.address > buf_.size() || need_bytes > buf_.size() - d.address || need_bytes > d.bytes)
                throw std::runtime_error("h5lite: dataset payload beyond end of file: " + name);
            p = bytes(d.address);
        }
        std::vector<double> out(cnt);
        for (std::uint64_t i = 0; i < cnt; i++) {
            unsigned char e[8] = {0};
            for (int b = 0; b < d.type_size; b++) e[b] = p[i * d.type_size + (d.big_endian ? d.type_size - 1 - b : b)];
            if (d.type_class == 1 && d.type_size == 4) { float v; std::memcpy(&v, e, 4); out[i] = v; }
            else if (d.type_class == 1 && d.type_size == 8) { double v; std::memcpy(&v, e, 8); out[i] = v; }
            else if (d.type_class == 0) {
                std::uint64_t u = 0;
                std::memcpy(&u, e, 8);
                if (d.is_signed && d.type_size < 8 && (u >> (8 * d.type_size - 1))) u |= ~0ull << (8 * d.type_size);
                out[i] = d.is_signed ? (double)(std::int64_t)u : (double)u;
            } else throw std::runtime_error("h5lite: unsupported element type in " + name);
        }
        return out;
    }

private:
    static constexpr std::uint64_t UNDEF = ~0ull;
    std::vector<char> buf_;
    std::map<std::string, H5Dataset> ds_;
    int so_ = 8, sl_ = 8;  // size of offsets / lengths
    long nodes_ = 0;       // B-tree nodes visited (cycle guard)
    std::uint64_t base_ = 0;

    const unsigned char *bytes(std::uint64_t off) const { return reinterpret_cast<const unsigned char *>(buf_.data()) + off; }
    void need(std::uint64_t off, std::uint64_t len) const {
        if (off > buf_.size() || len > buf_.size() - off) throw std::runtime_error("h5lite: truncated or corrupt file");
    }
    std::uint64_t le(std::uint64_t off, int n) const {
        need(off, n);
        std::uint64_t v = 0;
        for (int i = 0; i < n; i++) v |= (std::uint64_t)bytes(off)[i] << (8 * i);
        if (n < 8 && v == ((1ull << (8 * n)) - 1) && n == so_) return UNDEF;
        return v;
    }
    bool sig(std::uint64_t off, const char *s) const { need(off, 4); return std::memcmp(bytes(off), s, 4) == 0; }

    void parse() {
        static const unsigned char magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        need(0, 96);
        if (std::memcmp(bytes(0), magic, 8) != 0) throw std::runtime_error("h5lite: not an HDF5 file");
        const int version = bytes(0)[8];
        if (version != 0 && version != 1) throw std::runtime_error("h5lite: superblock version " + std::to_string(version) + " not supported (expected the v0 layout h5py writes by default)");
        so_ = bytes(0)[13];
        sl_ = bytes(0)[14];
        if (so_ != 8 && so_ != 4) throw std::runtime_error("h5lite: unsupported offset size");
        if (sl_ != 8 && sl_ != 4) throw std::runtime_error("h5lite: unsupported length size");
        std::uint64_t at = 24 + (version == 1 ? 4 : 0);
        base_ = le(at, so_);
        at += 4 * so_;  // base, free-space info, end of file, driver info
        // root group symbol table entry
        const std::uint64_t cache_type = le(at + 2 * so_, 4);
        if (cache_type != 1) throw std::runtime_error("h5lite: root group without cached symbol table");
        const std::uint64_t btree = le(at + 2 * so_ + 8, so_), heap = le(at + 2 * so_ + 8 + so_, so_);
        walk_tree(base_ + btree, heap_data(base_ + heap));
    }

    std::uint64_t heap_data(std::uint64_t heap) const {
        if (!sig(heap, "HEAP")) throw std::runtime_error("h5lite: bad local heap");
        return base_ + le(heap + 8 + 2 * sl_, so_);
    }

    void walk_tree(std::uint64_t node, std::uint64_t names, int depth = 0) {
        if (depth > 32 || ++nodes_ > 1000000) throw std::runtime_error("h5lite: B-tree too deep or cyclic");
        if (!sig(node, "TREE")) throw std::runtime_error("h5lite: bad B-tree node");
        const int type = bytes(node)[4], level = bytes(node)[5];
        const int used = (int)le(node + 6, 2);
        if (type != 0) throw std::runtime_error("h5lite: unexpected B-tree node type");
        std::uint64_t at = node + 8 + 2 * so_;  // first key
        for (int i = 0; i < used; i++) {
            at += sl_;  // key i
            const std::uint64_t child = base_ + le(at, so_);
            at += so_;
            if (level > 0) walk_tree(child, names, depth + 1);
            else symbol_node(child, names);
        }
    }

    void symbol_node(std::uint64_t node, std::uint64_t names) {
        if (!sig(node, "SNOD")) throw std::runtime_error("h5lite: bad symbol node");
        const int count = (int)le(node + 6, 2);
        std::uint64_t at = node + 8;
        for (int i = 0; i < count; i++, at += 2 * so_ + 24) {
            const std::uint64_t name_off = le(at, so_), header = le(at + so_, so_);
            const std::uint64_t cache_type = le(at + 2 * so_, 4);
            need(names + name_off, 1);
            const char *s = buf_.data() + names + name_off;
            const std::string name(s, strnlen(s, buf_.size() - (names + name_off)));
            if (cache_type == 1) continue;  // a sub-group: the weight files keep every tensor at the root
            H5Dataset d;
            if (object_header(base_ + header, d)) ds_[name] = d;
        }
    }

    // version-1 object header; returns false for objects that are not simple datasets
    bool object_header(std::uint64_t at, H5Dataset &d) {
        need(at, 16);
        if (bytes(at)[0] != 1) throw std::runtime_error("h5lite: object header version " + std::to_string(bytes(at)[0]) + " not supported");
        int remaining = (int)le(at + 2, 2);
        std::uint64_t block = at + 16, block_end = block + le(at + 8, 4);
        need(block, block_end - block);
        std::vector<std::pair<std::uint64_t, std::uint64_t>> more;
        bool have_space = false, have_type = false, have_layout = false;
        while (remaining > 0) {
            if (block + 8 > block_end) {
                if (more.empty()) break;
                block = more.back().first;
                block_end = block + more.back().second;
                more.pop_back();
                continue;
            }
            const int type = (int)le(block, 2), size = (int)le(block + 2, 2);
            const std::uint64_t body = block + 8;
            need(body, size);
            remaining--;
            if (type == 0x0001) {  // dataspace
                if (size < 4) throw std::runtime_error("h5lite: truncated dataspace message");
                const int ver = bytes(body)[0], rank = bytes(body)[1];
                if (rank > 32) throw std::runtime_error("h5lite: dataspace rank out of range");
                const std::uint64_t dims = body + (ver == 1 ? 8 : 4);
                d.dims.clear();
                for (int r = 0; r < rank; r++) d.dims.push_back(le(dims + (std::uint64_t)r * sl_, sl_));
                have_space = true;
            } else if (type == 0x0003) {  // datatype
                if (size < 8) throw std::runtime_error("h5lite: truncated datatype message");
                d.type_class = bytes(body)[0] & 0x0f;
                d.big_endian = (bytes(body)[1] & 1) != 0;
                d.is_signed = d.type_class == 0 ? (bytes(body)[1] & 8) != 0 : true;
                d.type_size = (int)le(body + 4, 4);
                have_type = true;
            } else if (type == 0x0008) {  // data layout
                const int ver = bytes(body)[0];
                if (ver != 3) throw std::runtime_error("h5lite: data layout version " + std::to_string(ver) + " not supported");
                const int cls = bytes(body)[1];
                if (cls == 1) {
                    d.address = le(body + 2, so_);
                    d.bytes = le(body + 2 + so_, sl_);
                    if (d.address != UNDEF) d.address += base_;
                } else if (cls == 0) {
                    const int sz = (int)le(body + 2, 2);
                    need(body + 4, sz);
                    d.compact.assign(bytes(body + 4), bytes(body + 4) + sz);
                } else throw std::runtime_error("h5lite: chunked datasets are not supported");
                have_layout = true;
            } else if (type == 0x0010) {  // continuation
                if (more.size() > 64) throw std::runtime_error("h5lite: too many header continuation blocks");
                const std::uint64_t cont = base_ + le(body, so_), clen = le(body + so_, sl_);
                need(cont, clen);
                more.push_back({cont, clen});
            }
            block = body + ((size + 7) & ~7);
        }
        if (!(have_space && have_type && have_layout)) return false;
        if (d.type_class != 0 && d.type_class != 1) return false;
        if (d.compact.empty() && d.address == UNDEF) d.dims.assign(1, 0);  // never written: no elements
        return true;
    }
};

// Same surface as the reference's LoadH5 (CrCNN/src/H5Easy.h:67-140) for the calls CnnBuilder makes.
class LoadH5 {
public:
    void setFileName(std::string name) { filename_ = name; file_.reset(); }
    void setVarName(std::string name) { variable_ = name; }
    std::vector<float> getDataVfloat() const {
        const H5File &f = open();
        auto it = f.datasets().find(variable_);
        if (it == f.datasets().end()) throw std::runtime_error("h5lite: no dataset named " + variable_ + " in " + filename_);
        if (it->second.type_class != 1) throw std::runtime_error(variable_ + " is not a float... you can't save this as a float.");  // H5Easy.cpp:603-607
        std::vector<double> v = f.read(variable_);
        return std::vector<float>(v.begin(), v.end());
    }
    std::vector<float> getData() const { return getDataVfloat(); }
    int getSize() const {
        const H5File &f = open();
        auto it = f.datasets().find(variable_);
        if (it == f.datasets().end()) throw std::runtime_error("h5lite: no dataset named " + variable_);
        return (int)it->second.count();
    }
    const H5File &file() const { return open(); }

private:
    const H5File &open() const {
        if (!file_) file_.reset(new H5File(filename_));
        return *file_;
    }
    std::string variable_, filename_;
    mutable std::shared_ptr<H5File> file_;
};

}  // namespace crcnn_b200
