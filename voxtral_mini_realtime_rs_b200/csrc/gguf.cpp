// gguf.cpp -- see gguf.h.  Behaviour (errors included) follows reference src/gguf/reader.rs.
#include "gguf.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <functional>
#include <cstdint>
#include <cstring>

#include "common.h"

namespace vox {

static const uint32_t kMagic = 0x46554747u;  // "GGUF" LE (reader.rs:13)
static const uint64_t kAlign = 32;           // reader.rs:14

// Product of the dims; saturates at UINT64_MAX on overflow (parse() rejects such tensors).
uint64_t GgufTensorInfo::num_elements() const {
    uint64_t n = 1;
    for (uint64_t d : dims) {
        if (d != 0 && n > UINT64_MAX / d) return UINT64_MAX;
        n *= d;
    }
    return n;
}

uint64_t GgufTensorInfo::byte_size() const {  // GgmlDtype::byte_size reader.rs:38-49
    uint64_t n = num_elements();
    if (n > UINT64_MAX / 4) return UINT64_MAX;
    switch (dtype) {
        case VOX_DTYPE_F32: return n * 4;
        case VOX_DTYPE_F16: return n * 2;
        default: return n / 32 * 18;
    }
}

std::vector<int64_t> GgufTensorInfo::shape() const {
    std::vector<int64_t> s(dims.rbegin(), dims.rend());
    return s;
}

Gguf::~Gguf() {
    if (map_base_) munmap(map_base_, map_len_);
}

Gguf *Gguf::open_file(const std::string &path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    VOX_CHECK(fd >= 0, VOX_EIO, "Failed to open %s", path.c_str());
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) {
        ::close(fd);
        fail(VOX_EIO, fmt("Failed to stat %s", path.c_str()));
    }
    void *p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    VOX_CHECK(p != MAP_FAILED, VOX_EIO, "Failed to mmap %s", path.c_str());
    Gguf *g = new Gguf();
    g->map_base_ = p;
    g->map_len_ = (size_t)st.st_size;
    g->shard_ptr_.push_back((const uint8_t *)p);
    g->shard_end_.push_back((uint64_t)st.st_size);
    g->total_len_ = (uint64_t)st.st_size;
    try {
        g->parse();
    } catch (...) {
        delete g;
        throw;
    }
    return g;
}

Gguf *Gguf::open_shards(const void *const *bufs, const size_t *lens, size_t n) {
    Gguf *g = new Gguf();
    uint64_t total = 0;
    for (size_t i = 0; i < n; ++i) {
        total += lens[i];
        g->shard_ptr_.push_back((const uint8_t *)bufs[i]);
        g->shard_end_.push_back(total);
    }
    g->total_len_ = total;
    try {
        g->parse();
    } catch (...) {
        delete g;
        throw;
    }
    return g;
}

// ShardedCursor::read (reader.rs:268-290): copy across shard boundaries.
void Gguf::read_at(uint64_t pos, void *dst, size_t n) const {
    // overflow-safe: `pos` may come straight from a (crafted) file
    VOX_CHECK(n <= total_len_ && pos <= total_len_ - n, VOX_EIO, "GGUF read past end (offset %llu + %zu > %llu)",
              (unsigned long long)pos, n, (unsigned long long)total_len_);
    uint8_t *out = (uint8_t *)dst;
    while (n > 0) {
        size_t si = std::upper_bound(shard_end_.begin(), shard_end_.end(), pos) - shard_end_.begin();
        uint64_t start = si ? shard_end_[si - 1] : 0;
        size_t local = (size_t)(pos - start);
        size_t avail = (size_t)(shard_end_[si] - start) - local;
        size_t take = std::min(avail, n);
        memcpy(out, shard_ptr_[si] + local, take);
        out += take;
        pos += take;
        n -= take;
    }
}

namespace {
struct Cursor {
    const Gguf *g;
    uint64_t pos;
    void (*rd)(const Gguf *, uint64_t, void *, size_t);
};
}  // namespace

void Gguf::parse() {
    uint64_t pos = 0;
    auto rd = [&](void *dst, size_t n, const char *what) {
        if (n > total_len_ || pos > total_len_ - n) fail(VOX_EIO, fmt("Failed to read %s", what));
        read_at(pos, dst, n);
        pos += n;
    };
    auto r_u32 = [&](const char *w) { uint32_t v; rd(&v, 4, w); return v; };
    auto r_u64 = [&](const char *w) { uint64_t v; rd(&v, 8, w); return v; };
    auto r_str = [&](const char *w) {
        uint64_t len = r_u64(w);
        VOX_CHECK(len <= total_len_, VOX_EIO, "Failed to read %s (bad string length)", w);
        std::string s((size_t)len, '\0');
        if (len) rd(&s[0], (size_t)len, w);
        return s;
    };
    uint32_t magic = r_u32("GGUF magic");
    VOX_CHECK(magic == kMagic, VOX_EFORMAT, "Invalid GGUF magic: 0x%08X (expected 0x%08X)", magic, kMagic);
    version_ = r_u32("GGUF version");
    VOX_CHECK(version_ == 2 || version_ == 3, VOX_EFORMAT, "Unsupported GGUF version: %u (expected 2 or 3)", version_);
    tensor_count_ = r_u64("tensor count");
    uint64_t kv_count = r_u64("metadata KV count");

    // skip_gguf_value (reader.rs:327-376); u32 values are additionally remembered.
    std::function<void(uint32_t, const std::string *, int)> skip;
    skip = [&](uint32_t t, const std::string *key, int depth) {
        VOX_CHECK(depth <= 16, VOX_EFORMAT, "GGUF metadata arrays nested deeper than 16");
        switch (t) {
            case 0: case 1: case 7: pos += 1; break;
            case 2: case 3: pos += 2; break;
            case 4: {
                uint32_t v = r_u32("metadata value");
                if (key) kv_u32_[*key] = v;
                break;
            }
            case 5: case 6: pos += 4; break;
            case 8: (void)r_str("metadata string"); break;
            case 9: {
                uint32_t et = r_u32("array type");
                uint64_t cnt = r_u64("array count");
                VOX_CHECK(cnt <= total_len_, VOX_EIO, "Failed to skip metadata value (array count %llu)", (unsigned long long)cnt);
                for (uint64_t i = 0; i < cnt; ++i) skip(et, nullptr, depth + 1);
                break;
            }
            case 10: case 11: case 12: pos += 8; break;
            default: fail(VOX_EFORMAT, fmt("Unknown GGUF metadata value type: %u", t));
        }
        VOX_CHECK(pos <= total_len_, VOX_EIO, "Failed to skip metadata value");
    };
    for (uint64_t i = 0; i < kv_count; ++i) {
        std::string key = r_str("metadata key");
        uint32_t vt = r_u32("metadata value type");
        skip(vt, &key, 0);
    }
    for (uint64_t i = 0; i < tensor_count_; ++i) {
        GgufTensorInfo t;
        t.name = r_str("tensor name");
        uint32_t nd = r_u32("ndims");
        VOX_CHECK(nd <= 8, VOX_EFORMAT, "Tensor '%s' has %u dims", t.name.c_str(), nd);
        for (uint32_t d = 0; d < nd; ++d) t.dims.push_back(r_u64("dim"));
        t.dtype = r_u32("dtype");
        VOX_CHECK(t.dtype <= 2, VOX_EFORMAT, "Unsupported GGML dtype code: %u", t.dtype);
        t.offset = r_u64("offset");
        names_.push_back(t.name);
        tensors_[t.name] = t;
    }
    data_offset_ = (pos + kAlign - 1) / kAlign * kAlign;  // reader.rs:177-179
    // every tensor's extent must lie inside the file (overflow-checked): the reference's reader errors out of
    // read_exact here; we refuse at parse time so that no later read can run past a shard
    VOX_CHECK(data_offset_ <= total_len_ || tensor_count_ == 0, VOX_EIO, "GGUF data section starts past the end of the file");
    for (const auto &kv : tensors_) {
        const GgufTensorInfo &t = kv.second;
        const uint64_t nb = t.byte_size();
        VOX_CHECK(t.num_elements() != UINT64_MAX && nb != UINT64_MAX, VOX_EFORMAT, "Tensor '%s': element count overflows", t.name.c_str());
        if (t.dtype == VOX_DTYPE_Q4_0)
            VOX_CHECK(t.num_elements() % 32 == 0, VOX_EFORMAT, "Tensor '%s': Q4_0 element count not a multiple of 32", t.name.c_str());
        const uint64_t room = total_len_ - data_offset_;
        VOX_CHECK(t.offset <= room && nb <= room - t.offset, VOX_EIO, "Tensor '%s' extends past the end of the file (offset %llu, %llu bytes)",
                  t.name.c_str(), (unsigned long long)t.offset, (unsigned long long)nb);
    }
}

const GgufTensorInfo *Gguf::find(const std::string &name) const {
    auto it = tensors_.find(name);
    return it == tensors_.end() ? nullptr : &it->second;
}

void Gguf::read_tensor(const GgufTensorInfo &t, void *dst) const {
    read_at(data_offset_ + t.offset, dst, (size_t)t.byte_size());
}

bool Gguf::kv_u32(const std::string &key, uint32_t *out) const {
    auto it = kv_u32_.find(key);
    if (it == kv_u32_.end()) return false;
    *out = it->second;
    return true;
}

}  // namespace vox
