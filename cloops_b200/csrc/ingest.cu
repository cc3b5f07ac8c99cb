// BEDPE ingest (cLoops/io.py:30-59 PET, :62-129 parseRawBedpe, :132-189 parseRawBedpe2, :192-203 txt2jd), host C++ behind
// the C ABI.  The reference reads the text line by line in one Python process (~0.25 M lines/s), writes "id cA cB" text per
// chromosome and parses that text again into the .jd matrix.  Here one reader thread inflates / reads the files into
// newline-aligned blocks and a pool of workers tokenizes them; every accepted cis PET lands in per-chromosome columns
// (cA, cB, opposite strands, line number) in file order, chromosomes in order of first appearance.
//
// Per line, in the reference's order (io.py:154-176):
//   split at tabs; skip when one field is "*" and one is "-1" (:159); skip when fewer than 6 fields (:161); PET(line) reads
//   fields 0-5, 8 and 9, so anything shorter than 10 fields raises and is skipped (:163-166); skip trans PETs (:168), PETs
//   outside the wanted chromosomes (:171), PETs with cB - cA < cut when cut > 0 (:174).  PET orients the two anchors by
//   start+end (:51-54) and takes floor((start+end)/2) as the centres (:55-56, python 2 integer division).
// Python's int() accepts more than plain decimal digits (blanks around the number, unicode digits, ...): a line with >= 10
// fields whose four coordinates are not all of the form [+-]?[0-9]{1,18} is not decided here but handed back to the caller
// with its line number ("odd" lines), which applies the reference's own expression to it.
// Lines end at "\n"; "\r\n" is read as "\n" (python 3 text mode, which the shimmed reference runs under).  A carriage return
// elsewhere would split the line under python 3 and not under python 2: such files are counted (bare_cr) and left to the
// caller's line-by-line reader.
#include <fcntl.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace {

struct Part {                      // the accepted PETs of one chromosome inside one block, file order
    std::string name;
    std::vector<int64_t> a, b;
    std::vector<int32_t> line;     // line index inside the block
    std::vector<uint8_t> opp;
};

struct BlockOut {
    std::vector<Part> parts;       // order of first appearance inside the block
    std::vector<std::pair<int32_t, std::string>> odd;
    int64_t n_lines = 0, bare_cr = 0;
};

struct Block {
    std::vector<char> buf;
    size_t len = 0;
    BlockOut* out = nullptr;
};

struct Wanted {
    std::vector<std::string> names;
    bool has(const char* s, size_t n) const {
        for (const std::string& c : names)
            if (c.size() == n && memcmp(c.data(), s, n) == 0) return true;
        return false;
    }
};

// [+-]?[0-9]{1,18}
inline bool plain_int(const char* s, const char* e, int64_t* out) {
    bool neg = false;
    if (s < e && (*s == '+' || *s == '-')) { neg = *s == '-'; ++s; }
    const ptrdiff_t nd = e - s;
    if (nd < 1 || nd > 18) return false;
    int64_t v = 0;
    for (; s < e; ++s) {
        const unsigned d = (unsigned)(*s - '0');
        if (d > 9) return false;
        v = v * 10 + d;
    }
    *out = neg ? -v : v;
    return true;
}

struct Parser {
    const Wanted& cs;
    const int64_t cut;
    BlockOut& out;
    int last = -1;                                         // Part of the previous accepted PET
    std::unordered_map<std::string, int> where;
    Parser(const Wanted& w, int64_t c, BlockOut& o) : cs(w), cut(c), out(o) {}

    Part& part_of(const char* s, size_t n) {
        if (last >= 0) {
            const std::string& ln = out.parts[last].name;
            if (ln.size() == n && memcmp(ln.data(), s, n) == 0) return out.parts[last];
        }
        std::string key(s, n);
        auto it = where.find(key);
        if (it == where.end()) {
            it = where.emplace(key, (int)out.parts.size()).first;
            out.parts.emplace_back();
            out.parts.back().name = key;
        }
        last = it->second;
        return out.parts[last];
    }

    void line(const char* p, const char* e, int32_t idx) {
        if (e > p && e[-1] == '\r') --e;                    // "\r\n"
        if (memchr(p, '\r', e - p)) { ++out.bare_cr; return; }
        const char *fs[10], *fe[10];
        int nf = 0;
        bool star = false, m1 = false;
        for (const char* q = p;;) {
            const char* t = (const char*)memchr(q, '\t', e - q);
            const char* fend = t ? t : e;
            if (nf < 10) { fs[nf] = q; fe[nf] = fend; }
            const ptrdiff_t len = fend - q;
            if (len == 1 && q[0] == '*') star = true;
            else if (len == 2 && q[0] == '-' && q[1] == '1') m1 = true;
            ++nf;
            if (!t) break;
            q = t + 1;
        }
        if (star && m1) return;                             // io.py:159
        if (nf < 10) return;                                // io.py:161 and the IndexError of PET(), :163-166
        int64_t sA, eA, sB, eB;
        if (!plain_int(fs[1], fe[1], &sA) || !plain_int(fs[2], fe[2], &eA) || !plain_int(fs[4], fe[4], &sB) ||
            !plain_int(fs[5], fe[5], &eB)) {
            out.odd.emplace_back(idx, std::string(p, e));
            return;
        }
        const size_t nA = fe[0] - fs[0];
        if (nA != (size_t)(fe[3] - fs[3]) || memcmp(fs[0], fs[3], nA) != 0) return;     // io.py:168
        if (!cs.names.empty() && !cs.has(fs[0], nA)) return;                             // io.py:171
        int64_t ta = sA + eA, tb = sB + eB;
        if (ta > tb) { const int64_t t = ta; ta = tb; tb = t; }                          // io.py:51-54
        const int64_t cA = ta >> 1, cB = tb >> 1;                                        // floor, io.py:55-56
        if (cut > 0 && cB - cA < cut) return;                                            // io.py:174
        const size_t n8 = fe[8] - fs[8];
        const bool opp = n8 != (size_t)(fe[9] - fs[9]) || memcmp(fs[8], fs[9], n8) != 0;
        Part& pt = part_of(fs[0], nA);
        pt.a.push_back(cA);
        pt.b.push_back(cB);
        pt.line.push_back(idx);
        pt.opp.push_back(opp ? 1 : 0);
    }

    void block(const char* p, size_t len) {
        const char* end = p + len;
        int32_t idx = 0;
        while (p < end) {
            const char* nl = (const char*)memchr(p, '\n', end - p);
            const char* e = nl ? nl : end;
            line(p, e, idx++);
            p = nl ? nl + 1 : end;
        }
        out.n_lines = idx;
    }
};

struct Queue {
    std::mutex m;
    std::condition_variable can_pop, can_push;
    std::deque<std::unique_ptr<Block>> q;
    size_t cap = 8;
    bool closed = false;
    void push(std::unique_ptr<Block> b) {
        std::unique_lock<std::mutex> l(m);
        can_push.wait(l, [&] { return q.size() < cap; });
        q.push_back(std::move(b));
        can_pop.notify_one();
    }
    std::unique_ptr<Block> pop() {
        std::unique_lock<std::mutex> l(m);
        can_pop.wait(l, [&] { return !q.empty() || closed; });
        if (q.empty()) return nullptr;
        std::unique_ptr<Block> b = std::move(q.front());
        q.pop_front();
        can_push.notify_one();
        return b;
    }
    void close() {
        std::lock_guard<std::mutex> l(m);
        closed = true;
        can_pop.notify_all();
    }
};

const size_t BLOCK_BYTES = 4u << 20;                        // 2^31 lines per block can not be reached

struct Source {                                             // plain file or gzip stream (io.py:148-151: by the ".gz" suffix)
    gzFile gz = nullptr;
    int fd = -1;
    bool open(const std::string& path) {
        if (path.size() >= 3 && path.compare(path.size() - 3, 3, ".gz") == 0) {
            gz = gzopen(path.c_str(), "rb");
            if (gz) gzbuffer(gz, 1u << 20);
            return gz != nullptr;
        }
        fd = ::open(path.c_str(), O_RDONLY);
        return fd >= 0;
    }
    long read(char* dst, size_t n) {                        // < 0: error, 0: end
        if (gz) return gzread(gz, dst, (unsigned)n);
        for (;;) {
            const ssize_t r = ::read(fd, dst, n);
            if (r < 0 && errno == EINTR) continue;
            return (long)r;
        }
    }
    ~Source() {
        if (gz) gzclose(gz);
        if (fd >= 0) ::close(fd);
    }
};

}  // namespace

struct cloops_bedpe {
    std::vector<std::string> names;
    std::vector<std::vector<int64_t>> a, b, line;
    std::vector<std::vector<uint8_t>> opp;
    std::vector<int64_t> odd_line;
    std::vector<std::string> odd_text;
    int64_t lines = 0, bare_cr = 0;
};

extern "C" int cloops_bedpe_parse(const char* const* paths, int n_paths, const char* const* chroms, int n_chroms, int64_t cut,
                                  int threads, cloops_bedpe** out) {
    using cloops::fail;
    if (!out || n_paths < 0 || n_chroms < 0 || (n_paths > 0 && !paths) || (n_chroms > 0 && !chroms))
        return fail(CLOOPS_EINVAL, "bad argument");
    *out = nullptr;
    Wanted cs;
    for (int k = 0; k < n_chroms; ++k) cs.names.emplace_back(chroms[k]);
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    if (threads > 64) threads = 64;

    // one slot per block, in file order; references into a deque stay valid while it grows
    std::vector<std::deque<BlockOut>> results((size_t)n_paths);
    Queue queue;
    queue.cap = 2 * (size_t)threads + 2;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            while (std::unique_ptr<Block> b = queue.pop()) {
                Parser ps(cs, cut, *b->out);
                ps.block(b->buf.data(), b->len);
            }
        });

    const bool trace = getenv("CLOOPS_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    std::vector<std::string> errs((size_t)n_paths);
    auto read_file = [&](int f) {
        Source src;
        if (!src.open(paths[f])) { errs[f] = std::string("cannot open ") + paths[f]; return; }
        std::vector<char> carry;                            // the unfinished last line of the previous read
        for (bool eof = false; !eof;) {
            std::unique_ptr<Block> b(new Block());
            b->buf.resize(carry.size() + BLOCK_BYTES);
            if (!carry.empty()) memcpy(b->buf.data(), carry.data(), carry.size());
            size_t have = carry.size();
            carry.clear();
            while (have < b->buf.size()) {
                const long r = src.read(b->buf.data() + have, b->buf.size() - have);
                if (r < 0) { errs[f] = std::string("read error in ") + paths[f]; return; }
                if (r == 0) { eof = true; break; }
                have += (size_t)r;
            }
            size_t len = have;
            if (!eof) {                                     // cut at the last newline; the rest opens the next block
                const char* base = b->buf.data();
                const void* nl = memrchr(base, '\n', have);
                len = nl ? (size_t)((const char*)nl - base) + 1 : 0;
                carry.assign(base + len, base + have);
            }
            if (len == 0) continue;                         // one line longer than a block: keep reading
            b->len = len;
            results[f].emplace_back();
            b->out = &results[f].back();
            queue.push(std::move(b));
        }
    };
    {
        // the replicate files are read (and inflated: one gzip stream is serial) side by side
        std::atomic<int> next(0);
        std::vector<std::thread> readers;
        const int n_readers = n_paths < threads ? n_paths : threads;
        for (int t = 0; t < n_readers; ++t)
            readers.emplace_back([&] {
                for (int f; (f = next.fetch_add(1)) < n_paths;) read_file(f);
            });
        for (std::thread& t : readers) t.join();
    }
    const auto t_read = std::chrono::steady_clock::now();
    queue.close();
    for (std::thread& t : pool) t.join();
    const auto t_join = std::chrono::steady_clock::now();
    for (const std::string& e : errs)
        if (!e.empty()) return fail(CLOOPS_EINVAL, "%s", e.c_str());

    std::unique_ptr<cloops_bedpe> h(new cloops_bedpe());
    std::unordered_map<std::string, int> where;
    std::vector<size_t> total;
    std::vector<BlockOut*> blocks;                          // command-line order of the files, file order inside
    for (std::deque<BlockOut>& file : results)
        for (BlockOut& r : file) blocks.push_back(&r);
    for (const BlockOut* r : blocks)
        for (const Part& pt : r->parts) {
            auto it = where.find(pt.name);
            if (it == where.end()) {
                it = where.emplace(pt.name, (int)h->names.size()).first;
                h->names.push_back(pt.name);
                total.push_back(0);
            }
            total[it->second] += pt.a.size();
        }
    const size_t nc = h->names.size();
    h->a.resize(nc); h->b.resize(nc); h->line.resize(nc); h->opp.resize(nc);
    for (size_t c = 0; c < nc; ++c) {
        h->a[c].reserve(total[c]); h->b[c].reserve(total[c]); h->line[c].reserve(total[c]); h->opp[c].reserve(total[c]);
    }
    int64_t base = 0;
    for (BlockOut* rp : blocks) {
        BlockOut& r = *rp;
        for (Part& pt : r.parts) {
            const int c = where[pt.name];
            h->a[c].insert(h->a[c].end(), pt.a.begin(), pt.a.end());
            h->b[c].insert(h->b[c].end(), pt.b.begin(), pt.b.end());
            h->opp[c].insert(h->opp[c].end(), pt.opp.begin(), pt.opp.end());
            for (int32_t l : pt.line) h->line[c].push_back(base + l);
            std::vector<int64_t>().swap(pt.a);
            std::vector<int64_t>().swap(pt.b);
            std::vector<int32_t>().swap(pt.line);
            std::vector<uint8_t>().swap(pt.opp);
        }
        for (auto& o : r.odd) {
            h->odd_line.push_back(base + o.first);
            h->odd_text.push_back(std::move(o.second));
        }
        base += r.n_lines;
        h->bare_cr += r.bare_cr;
    }
    h->lines = base;
    if (trace) {
        const auto t_end = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[cloops] bedpe ingest: %lld lines, %d tokenizers: read+queue %.1f ms, drain %.1f ms, stitch %.1f ms\n",
                (long long)base, threads, ms(t_start, t_read), ms(t_read, t_join), ms(t_join, t_end));
    }
    *out = h.release();
    return 0;
}

extern "C" int64_t cloops_bedpe_lines(const cloops_bedpe* h) { return h ? h->lines : -1; }
extern "C" int64_t cloops_bedpe_bare_cr(const cloops_bedpe* h) { return h ? h->bare_cr : -1; }
extern "C" int cloops_bedpe_n_chroms(const cloops_bedpe* h) { return h ? (int)h->names.size() : -1; }

extern "C" const char* cloops_bedpe_chrom(const cloops_bedpe* h, int k, int64_t* name_len, int64_t* n_pets) {
    if (!h || k < 0 || k >= (int)h->names.size()) return nullptr;
    if (name_len) *name_len = (int64_t)h->names[k].size();
    if (n_pets) *n_pets = (int64_t)h->a[k].size();
    return h->names[k].data();
}

extern "C" int cloops_bedpe_fetch(const cloops_bedpe* h, int k, int64_t* cA, int64_t* cB, uint8_t* opposite, int64_t* line_no) {
    using cloops::fail;
    if (!h || k < 0 || k >= (int)h->names.size()) return fail(CLOOPS_EINVAL, "no such chromosome");
    const size_t n = h->a[k].size();
    if (cA && n) memcpy(cA, h->a[k].data(), n * sizeof(int64_t));
    if (cB && n) memcpy(cB, h->b[k].data(), n * sizeof(int64_t));
    if (opposite && n) memcpy(opposite, h->opp[k].data(), n);
    if (line_no && n) memcpy(line_no, h->line[k].data(), n * sizeof(int64_t));
    return 0;
}

extern "C" int64_t cloops_bedpe_n_odd(const cloops_bedpe* h) { return h ? (int64_t)h->odd_line.size() : -1; }

extern "C" const char* cloops_bedpe_odd(const cloops_bedpe* h, int64_t k, int64_t* line_no, int64_t* len) {
    if (!h || k < 0 || k >= (int64_t)h->odd_line.size()) return nullptr;
    if (line_no) *line_no = h->odd_line[k];
    if (len) *len = (int64_t)h->odd_text[k].size();
    return h->odd_text[k].data();
}

extern "C" void cloops_bedpe_free(cloops_bedpe* h) { delete h; }
