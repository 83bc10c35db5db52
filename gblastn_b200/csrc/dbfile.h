// dbfile.h — BLAST database volume files (format version 4, nucleotide): the .nin index and the .nsq
// packed-sequence file, read on the host and streamed to HBM as they are.
//
// Layout (what the reference's reader parses, objtools/blast/seqdb_reader/seqdbfile.cpp:195-250,
// impl/seqdbfile.hpp:786-795, seqdbvol.cpp:263-285):
//   .nin  big-endian Uint4 version (4), Uint4 seqtype (0 = nucleotide), length-prefixed title, length-
//         prefixed create date (NUL-padded by the writer), Uint4 nseq, Uint8 total length (LITTLE-endian,
//         "SeqDB_GetBroken"), Uint4 max length, then three arrays of nseq + 1 big-endian Uint4 offsets:
//         header (.nhr), sequence (.nsq), ambiguity (.nsq).
//   .nsq  one NUL byte, then per sequence its ncbi2na bytes [seq[i], amb[i]) — the last byte holds the
//         remaining 0-3 bases in its high bits and their count in its low two bits — followed by the
//         ambiguity words [amb[i], seq[i+1]) that only the traceback stage uses.
// The preliminary search reads exactly the bytes [seq[i], amb[i]) (api/seqsrc_seqdb.cpp:283-388), so
// the whole .nsq goes to the device unchanged and sequence i starts at byte seq[i].
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace bn {

struct DbIndex {
    int32_t version = 0, seqtype = 0, n_seq = 0, max_len = 0;
    uint64_t total_len = 0;
    std::string title, date;
    std::vector<uint32_t> hdr_off, seq_off, amb_off;   // n_seq + 1 each
};

// Parses a .nin file.  Returns false with a message on any malformation.
bool read_nin(const char *path, DbIndex &out, std::string &err);

// Read-only memory map of a file.
class MappedFile {
public:
    MappedFile() = default;
    ~MappedFile();
    MappedFile(const MappedFile &) = delete;
    MappedFile &operator=(const MappedFile &) = delete;
    bool open(const char *path, std::string &err);
    const uint8_t *data() const { return data_; }
    int64_t size() const { return size_; }
private:
    const uint8_t *data_ = nullptr;
    int64_t size_ = 0;
};

// Per-sequence byte offsets into the .nsq and lengths in bases (the count of bases in a sequence's
// last byte sits in that byte).  Verifies total / max length against the header.
bool sequence_table(const DbIndex &idx, const uint8_t *nsq, int64_t nsq_bytes,
                    std::vector<int64_t> &byte_off, std::vector<int32_t> &seq_len, std::string &err);

// Ambiguity data of a nucleotide volume: the words [amb[i], seq[i+1]) of the .nsq (big-endian Int4; the first is the
// number of entries, its top bit selecting the 8-byte "new" entry form) decoded like s_SeqDBRebuildDNA_NA8
// (objtools/blast/seqdb_reader/seqdbvol.cpp:832-870, accessors :640-745) into runs {first base, number of bases,
// blastna code}: the ncbi4na residue of an entry mapped through SeqDB_ncbina8_to_blastna8 (:561-578).
// first[i] .. first[i + 1] index the runs of sequence i (n_seq + 1 entries); runs are flat triples in file order
// (the reference applies them in that order, later entries overwrite earlier ones).
bool ambiguity_table(const DbIndex &idx, const uint8_t *nsq, int64_t nsq_bytes, std::vector<int64_t> &first,
                     std::vector<int32_t> &runs, std::string &err);

// Writes a volume (our in-memory layout: sequence i = (seq_len[i] + 3) / 4 bytes at seq_byte_off[i])
// as .nin + .nsq without ambiguity data or deflines.
bool write_volume(const char *nin_path, const char *nsq_path, const char *title, const uint8_t *packed,
                  const int64_t *seq_byte_off, const int32_t *seq_len, int32_t n_seq, std::string &err);

}  // namespace bn
