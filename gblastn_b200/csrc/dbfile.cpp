// dbfile.cpp — see dbfile.h.
#include "dbfile.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace bn {

namespace {

struct File {
    FILE *f = nullptr;
    explicit File(const char *path, const char *mode) { f = fopen(path, mode); }
    ~File() { if (f) fclose(f); }
};

bool read_exact(FILE *f, void *dst, size_t n) { return fread(dst, 1, n, f) == n; }

bool read_be32(FILE *f, uint32_t &v)
{
    uint8_t b[4];
    if (!read_exact(f, b, 4)) return false;
    v = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    return true;
}

bool read_string(FILE *f, std::string &s)
{
    uint32_t len;
    if (!read_be32(f, len) || len > (1u << 20)) return false;
    s.resize(len);
    return len == 0 || read_exact(f, &s[0], len);
}

void put_be32(std::vector<uint8_t> &o, uint32_t v)
{
    o.push_back((uint8_t)(v >> 24)); o.push_back((uint8_t)(v >> 16)); o.push_back((uint8_t)(v >> 8)); o.push_back((uint8_t)v);
}

}  // namespace

bool read_nin(const char *path, DbIndex &out, std::string &err)
{
    File fh(path, "rb");
    if (!fh.f) { err = std::string("cannot open ") + path; return false; }
    uint32_t v;
    if (!read_be32(fh.f, v)) { err = "truncated index file"; return false; }
    out.version = (int32_t)v;
    if (out.version != 4) { err = "not a version 4 BLAST database index"; return false; }
    if (!read_be32(fh.f, v)) { err = "truncated index file"; return false; }
    out.seqtype = (int32_t)v;
    if (out.seqtype == 1) { err = "protein database: the blastn path needs a nucleotide volume"; return false; }
    if (!read_string(fh.f, out.title) || !read_string(fh.f, out.date)) { err = "truncated index file (title/date)"; return false; }
    if (!read_be32(fh.f, v)) { err = "truncated index file"; return false; }
    out.n_seq = (int32_t)v;
    uint8_t b8[8];
    if (!read_exact(fh.f, b8, 8)) { err = "truncated index file"; return false; }
    out.total_len = 0;
    for (int i = 7; i >= 0; i--) out.total_len = (out.total_len << 8) | b8[i];      // little-endian
    if (!read_be32(fh.f, v)) { err = "truncated index file"; return false; }
    out.max_len = (int32_t)v;
    if (out.n_seq < 0) { err = "negative sequence count"; return false; }
    const size_t n = (size_t)out.n_seq + 1;
    std::vector<uint8_t> raw(4 * n);
    std::vector<uint32_t> *arrs[3] = {&out.hdr_off, &out.seq_off, &out.amb_off};
    for (auto *a : arrs) {
        if (!read_exact(fh.f, raw.data(), raw.size())) { err = "truncated index file (offset arrays)"; return false; }
        a->resize(n);
        for (size_t i = 0; i < n; i++)
            (*a)[i] = ((uint32_t)raw[4 * i] << 24) | ((uint32_t)raw[4 * i + 1] << 16) | ((uint32_t)raw[4 * i + 2] << 8) | raw[4 * i + 3];
    }
    return true;
}

MappedFile::~MappedFile()
{
    if (data_ && size_ > 0) munmap(const_cast<uint8_t *>(data_), (size_t)size_);
}

bool MappedFile::open(const char *path, std::string &err)
{
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) { err = std::string("cannot open ") + path; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); err = std::string("cannot stat ") + path; return false; }
    void *p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { err = std::string("cannot map ") + path; return false; }
    data_ = static_cast<const uint8_t *>(p);
    size_ = (int64_t)st.st_size;
    return true;
}

bool sequence_table(const DbIndex &idx, const uint8_t *nsq, int64_t nsq_bytes, std::vector<int64_t> &byte_off,
                    std::vector<int32_t> &seq_len, std::string &err)
{
    byte_off.resize((size_t)idx.n_seq);
    seq_len.resize((size_t)idx.n_seq);
    uint64_t total = 0;
    int32_t longest = 0;
    for (int32_t i = 0; i < idx.n_seq; i++) {
        const int64_t start = idx.seq_off[(size_t)i], end = idx.amb_off[(size_t)i];     // GetSeqStartEnd, nucleotide
        if (start < 1 || end <= start || end > nsq_bytes || (int64_t)idx.seq_off[(size_t)i + 1] < end) {
            err = "sequence offsets of the index do not fit the sequence file";
            return false;
        }
        const int64_t len = (end - start - 1) * 4 + (nsq[end - 1] & 3);                  // GetSeqLengthExact
        if (len > INT32_MAX) { err = "sequence longer than 2^31 bases"; return false; }
        byte_off[(size_t)i] = start;
        seq_len[(size_t)i] = (int32_t)len;
        total += (uint64_t)len;
        longest = std::max(longest, (int32_t)len);
    }
    if (total != idx.total_len || longest != idx.max_len) {
        err = "sequence lengths disagree with the total / maximum length of the index header";
        return false;
    }
    return true;
}

bool ambiguity_table(const DbIndex &idx, const uint8_t *nsq, int64_t nsq_bytes, std::vector<int64_t> &first,
                     std::vector<int32_t> &runs, std::string &err)
{
    static const int32_t kNcbi4naToBlastna[16] = {15, 0, 1, 6, 2, 4, 9, 13, 3, 8, 5, 12, 7, 11, 10, 14};
    first.assign((size_t)idx.n_seq + 1, 0);
    runs.clear();
    auto word = [&](int64_t at) -> uint32_t {
        return ((uint32_t)nsq[at] << 24) | ((uint32_t)nsq[at + 1] << 16) | ((uint32_t)nsq[at + 2] << 8) | nsq[at + 3];
    };
    for (int32_t i = 0; i < idx.n_seq; i++) {
        const int64_t a = idx.amb_off[(size_t)i], e = idx.seq_off[(size_t)i + 1];
        first[(size_t)i] = (int64_t)(runs.size() / 3);
        if (e <= a) continue;
        if (e > nsq_bytes || (e - a) % 4 != 0 || e - a < 4) { err = "malformed ambiguity data"; return false; }
        const int64_t n_words = (e - a) / 4;
        uint32_t n = word(a);
        const bool wide = (n & 0x80000000u) != 0;
        n &= 0x7FFFFFFFu;
        if ((int64_t)n > n_words - 1) { err = "ambiguity entry count exceeds the data"; return false; }
        for (uint32_t k = 1; k < n + 1; k++) {
            const uint32_t w = word(a + 4 * (int64_t)k);
            const int32_t code = kNcbi4naToBlastna[(w >> 28) & 0xF];
            int32_t len, pos;
            if (wide) {
                if ((int64_t)k + 1 > n_words - 1) { err = "truncated ambiguity entry"; return false; }
                len = (int32_t)((w >> 16) & 0xFFF) + 1;
                pos = (int32_t)word(a + 4 * (int64_t)(k + 1));
                ++k;
            } else {
                len = (int32_t)((w >> 24) & 0xF) + 1;
                pos = (int32_t)(w & 0xFFFFFF);
            }
            runs.push_back(pos); runs.push_back(len); runs.push_back(code);
        }
    }
    first[(size_t)idx.n_seq] = (int64_t)(runs.size() / 3);
    return true;
}

bool write_volume(const char *nin_path, const char *nsq_path, const char *title, const uint8_t *packed,
                  const int64_t *seq_byte_off, const int32_t *seq_len, int32_t n_seq, std::string &err)
{
    File sq(nsq_path, "wb");
    if (!sq.f) { err = std::string("cannot create ") + nsq_path; return false; }
    std::vector<uint32_t> seq_off((size_t)n_seq + 1), amb_off((size_t)n_seq + 1);
    uint64_t pos = 1, total = 0;
    int32_t longest = 0;
    fputc(0, sq.f);
    for (int32_t i = 0; i < n_seq; i++) {
        const int32_t len = seq_len[i], whole = len / 4, rem = len & 3;
        const uint8_t *src = packed + seq_byte_off[i];
        seq_off[(size_t)i] = (uint32_t)pos;
        if (whole && fwrite(src, 1, (size_t)whole, sq.f) != (size_t)whole) { err = "write failed"; return false; }
        const uint8_t last = rem ? (uint8_t)((src[whole] & (uint8_t)(0xFF << (8 - 2 * rem))) | rem) : 0;
        fputc(last, sq.f);
        pos += (uint64_t)whole + 1;
        amb_off[(size_t)i] = (uint32_t)pos;           // no ambiguity data
        if (pos > 0xFFFFFFFFull) { err = "volume exceeds the 4 GB offset range of a version 4 index"; return false; }
        total += (uint64_t)len;
        longest = std::max(longest, len);
    }
    seq_off[(size_t)n_seq] = (uint32_t)pos; amb_off[(size_t)n_seq] = (uint32_t)pos;

    std::vector<uint8_t> o;
    put_be32(o, 4); put_be32(o, 0);
    const std::string t = title ? title : "";
    put_be32(o, (uint32_t)t.size()); o.insert(o.end(), t.begin(), t.end());
    std::string date = "Jan 1, 2000  12:00 AM";
    while ((o.size() + 4 + date.size()) % 8 != 0) date.push_back('\0');     // the fields that follow stay aligned
    put_be32(o, (uint32_t)date.size()); o.insert(o.end(), date.begin(), date.end());
    put_be32(o, (uint32_t)n_seq);
    for (int i = 0; i < 8; i++) o.push_back((uint8_t)(total >> (8 * i)));
    put_be32(o, (uint32_t)longest);
    for (int32_t i = 0; i <= n_seq; i++) put_be32(o, 0);                     // no .nhr
    for (uint32_t v : seq_off) put_be32(o, v);
    for (uint32_t v : amb_off) put_be32(o, v);
    File ix(nin_path, "wb");
    if (!ix.f) { err = std::string("cannot create ") + nin_path; return false; }
    if (fwrite(o.data(), 1, o.size(), ix.f) != o.size()) { err = "write failed"; return false; }
    return true;
}

}  // namespace bn
