// tests/native/host_sanitizer_driver.cpp -- AddressSanitizer / UBSan exercise of the host-side C++ behind the C ABI
// (GGUF reader over files and shards, Tekken tokenizer with its own JSON / base64 / UTF-8 code, audio plumbing):
// well-formed inputs, truncations at many offsets and byte flips.  Malformed inputs must be rejected with vox::Error or
// decoded harmlessly -- never a memory error.  Built and run by tests/test_host_sanitizers.py (no GPU).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "common.h"
#include "gguf.h"
#include "tokenizer.h"
#include "audio_host.h"
using namespace vox;
static std::string slurp(const char* p){ std::ifstream f(p, std::ios::binary); std::stringstream ss; ss<<f.rdbuf(); return ss.str(); }
int main(int argc, char** argv){
    if (argc < 3) { fprintf(stderr, "usage: drv tiny.gguf tekken.json\n"); return 2; }
    const char* gguf_path = argv[1]; const char* tok_path = argv[2];
    // ---- GGUF: file, shards, every tensor, truncated / corrupted inputs
    Gguf* g = Gguf::open_file(gguf_path);
    size_t bytes=0; for (auto& n : g->names()){ const GgufTensorInfo* t=g->find(n); std::vector<uint8_t> buf(t->byte_size()); g->read_tensor(*t, buf.data()); bytes+=buf.size(); (void)t->shape(); }
    printf("gguf v%u tensors %llu bytes %zu\n", g->version(), (unsigned long long)g->tensor_count(), bytes);
    delete g;
    std::string all = slurp(gguf_path);
    { // three shards at odd boundaries
        size_t a = all.size()/3+1, b = 2*all.size()/3+5;
        const void* bufs[3] = {all.data(), all.data()+a, all.data()+b}; size_t lens[3] = {a, b-a, all.size()-b};
        Gguf* s = Gguf::open_shards(bufs, lens, 3);
        for (auto& n : s->names()){ const GgufTensorInfo* t=s->find(n); std::vector<uint8_t> buf(t->byte_size()); s->read_tensor(*t, buf.data()); }
        delete s;
    }
    int rejected=0, accepted=0;
    for (size_t cut : {size_t(0), size_t(3), size_t(4), size_t(11), size_t(24), size_t(100), size_t(1000), all.size()/2, all.size()-1}) {
        const void* bufs[1] = {all.data()}; size_t lens[1] = {cut};
        try { Gguf* s = Gguf::open_shards(bufs, lens, 1); for (auto& n : s->names()){ const GgufTensorInfo* t=s->find(n); std::vector<uint8_t> buf(t->byte_size()); try { s->read_tensor(*t, buf.data()); } catch (const Error&) {} } delete s; ++accepted; }
        catch (const Error&) { ++rejected; }
    }
    // bit flips in the header / index region
    for (size_t pos = 0; pos < 4096 && pos < all.size(); pos += 7) {
        std::string m = all; m[pos] = (char)(m[pos] ^ 0xA5);
        const void* bufs[1] = {m.data()}; size_t lens[1] = {m.size()};
        try { Gguf* s = Gguf::open_shards(bufs, lens, 1); delete s; ++accepted; } catch (const Error&) { ++rejected; } catch (const std::exception&) { ++rejected; }
    }
    printf("gguf fuzz: %d rejected, %d accepted\n", rejected, accepted);
    { // crafted index entries: offsets near 2^64, overflowing dims, deeply nested metadata arrays
        auto put = [](std::string& b, const void* v, size_t n){ b.append((const char*)v, n); };
        auto u32 = [&](std::string& b, uint32_t v){ put(b, &v, 4); };
        auto u64 = [&](std::string& b, uint64_t v){ put(b, &v, 8); };
        auto str = [&](std::string& b, const char* s){ u64(b, strlen(s)); b.append(s); };
        auto craft = [&](uint64_t off, uint64_t d0, uint64_t d1, uint32_t dtype, int nest){
            std::string b; u32(b, 0x46554747u); u32(b, 3); u64(b, 1); u64(b, nest ? 1 : 0);
            if (nest) { str(b, "k"); u32(b, 9); for (int i = 0; i < nest; ++i) { u32(b, 9); u64(b, 1); } u32(b, 4); u64(b, 1); u32(b, 7); }
            str(b, "t"); u32(b, 2); u64(b, d0); u64(b, d1); u32(b, dtype); u64(b, off);
            b.append(256, '\0');
            return b;
        };
        int bad_ok = 0, bad_rej = 0;
        const uint64_t M = ~0ull;
        std::vector<std::string> cases = { craft(M - 8, 4, 4, 0, 0), craft(M, 32, 1, 2, 0), craft(0, M / 2, 4, 0, 0), craft(0, 1ull << 33, 1ull << 33, 1, 0),
                                           craft(64, 1ull << 62, 8, 2, 0), craft(0, 4, 4, 0, 1000), craft(1ull << 63, 4, 4, 0, 0), craft(0, 33, 1, 2, 0) };
        for (auto& c : cases) {
            const void* bufs[1] = {c.data()}; size_t lens[1] = {c.size()};
            try { Gguf* s = Gguf::open_shards(bufs, lens, 1);
                  for (auto& n : s->names()){ const GgufTensorInfo* t=s->find(n); std::vector<uint8_t> buf((size_t)std::min<uint64_t>(t->byte_size(), 1u << 20)); s->read_tensor(*t, buf.data()); }
                  delete s; ++bad_ok; }
            catch (const Error&) { ++bad_rej; }
        }
        printf("gguf crafted: %d rejected, %d accepted\n", bad_rej, bad_ok);
        if (bad_ok != 0) { fprintf(stderr, "crafted GGUF accepted\n"); return 1; }
    }
    { // tokenizer: deep nesting and an unterminated trailing number must be errors, not crashes
        std::string deep(100000, '['); int rej = 0;
        try { Tokenizer* x = Tokenizer::from_json(deep.data(), deep.size()); delete x; } catch (const Error&) { ++rej; } catch (const std::exception&) { ++rej; }
        std::vector<char> num = {'{','"','a','"',':','1','2','3','4','5'};   // exactly-sized heap buffer, no terminator
        try { Tokenizer* x = Tokenizer::from_json(num.data(), num.size()); delete x; } catch (const Error&) { ++rej; } catch (const std::exception&) { ++rej; }
        printf("tokenizer crafted: %d rejected\n", rej);
        if (rej != 2) return 1;
    }
    // ---- tokenizer: good file, truncations, byte flips
    std::string tj = slurp(tok_path);
    Tokenizer* t = Tokenizer::from_json(tj.data(), tj.size());
    std::vector<uint32_t> ids; for (uint32_t i = 0; i < 1700; ++i) ids.push_back(i);
    std::string text = t->decode(ids.data(), ids.size());
    printf("tokenizer vocab %zu decoded %zu bytes\n", t->vocab_size(), text.size());
    delete t;
    int trej=0, tacc=0;
    for (size_t cut = 0; cut < tj.size(); cut += tj.size()/97 + 1) {
        try { Tokenizer* x = Tokenizer::from_json(tj.data(), cut); std::string d = x->decode(ids.data(), 50); delete x; ++tacc; } catch (const Error&) { ++trej; } catch (const std::exception&) { ++trej; }
    }
    for (size_t pos = 0; pos < tj.size(); pos += tj.size()/211 + 1) {
        std::string m = tj; m[pos] = (char)(m[pos] ^ 0x5A);
        try { Tokenizer* x = Tokenizer::from_json(m.data(), m.size()); std::string d = x->decode(ids.data(), 200); delete x; ++tacc; } catch (const Error&) { ++trej; } catch (const std::exception&) { ++trej; }
    }
    printf("tokenizer fuzz: %d rejected, %d accepted\n", trej, tacc);
    // ---- audio plumbing
    for (size_t n : {size_t(0), size_t(1), size_t(159), size_t(160), size_t(1279), size_t(1280), size_t(256000), size_t(480001)}) {
        int64_t out[5]; stream_progress(n, false, 4, 38, out); stream_progress(n, true, 4, 38, out);
        auto plan = chunk_plan(n, 1500, 0); (void)plan;
    }
    std::vector<float> te(3072); time_embedding(6.0f, 3072, te.data());
    printf("done\n");
    return 0;
}
