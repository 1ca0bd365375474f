// db_io.cu -- the step before the hot path: FASTA -> (residues, offsets), and an on-disk
// form of the packed, length-sorted database so that a scan service does not sort and pack at every start.
//
//   psb_fasta_read / psb_fasta_free      plain or multi-line FASTA -> concatenated residues + offsets + names
//   psb_db_from_fasta                    FASTA -> resident packed database (psb_db_create on the parsed arrays)
//   psb_db_save / psb_db_load            the packed database as one file:
//
//       offset  size              field
//       0       8                 magic "PSBDB\0\1\0"
//       8       4                 bits per residue (2, 3 or 5)
//       12      4                 alphabet size of the matrix the residues were mapped with
//       16      8                 n          subjects
//       24      8                 residues
//       32      8                 words      32-bit words of packed residues
//       40      256               mapper     byte -> matrix column used when packing
//       296     8                 (reserved, zero)
//       304     4*n               len        sorted order (length descending, stable)
//       ..      4*n               perm       sorted position -> original subject id
//       ..      8*(n+1)           word_off   first word of each sorted subject
//       ..      4*words           words      residues, first residue in the low bits, each subject on a fresh word
//
//   The reference has no counterpart (callers hand &[u8] to every call, [REF src/aligner/mod.rs:397-452]).
#include "psb_db.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "psb_internal.h"

using namespace psb;

struct psb_fasta {
    int64_t n = 0;
    std::vector<uint8_t> cat;
    std::vector<int64_t> off;
    std::vector<char> names;
    std::vector<int64_t> name_off;
};

extern "C" {

psb_fasta_t *psb_fasta_read(const char *path) {
    if (!path) { set_error("psb_fasta_read: NULL path"); return nullptr; }
    std::FILE *f = std::fopen(path, "rb");
    if (!f) { set_error(std::string("psb_fasta_read: cannot open ") + path); return nullptr; }
    psb_fasta *fa = new psb_fasta();
    fa->off.push_back(0);
    fa->name_off.push_back(0);
    std::vector<char> buf(1 << 20);
    bool in_header = false, at_line_start = true, have_record = false;
    size_t got;
    while ((got = std::fread(buf.data(), 1, buf.size(), f)) > 0) {
        for (size_t x = 0; x < got; ++x) {
            const char ch = buf[x];
            if (in_header) {
                if (ch == '\n') { in_header = false; at_line_start = true; fa->name_off.push_back((int64_t)fa->names.size()); }
                else if (ch != '\r') fa->names.push_back(ch);
                continue;
            }
            if (ch == '\n') { at_line_start = true; continue; }
            if (at_line_start && ch == '>') {
                if (have_record) fa->off.push_back((int64_t)fa->cat.size());
                have_record = true; in_header = true;
                continue;
            }
            at_line_start = false;
            if (ch == '\r' || ch == ' ' || ch == '\t' || ch == '*') continue;   // a trailing '*' is the stop symbol of protein FASTA
            if (!have_record) { have_record = true; fa->name_off.push_back(0); }   // headerless file: one anonymous record
            fa->cat.push_back((uint8_t)ch);
        }
    }
    std::fclose(f);
    if (in_header) fa->name_off.push_back((int64_t)fa->names.size());
    if (have_record) fa->off.push_back((int64_t)fa->cat.size());
    fa->n = (int64_t)fa->off.size() - 1;
    while ((int64_t)fa->name_off.size() < fa->n + 1) fa->name_off.push_back((int64_t)fa->names.size());
    if (fa->n <= 0) { set_error(std::string("psb_fasta_read: no sequence in ") + path); delete fa; return nullptr; }
    return fa;
}
int64_t psb_fasta_count(const psb_fasta_t *fa) { return fa ? fa->n : 0; }
const uint8_t *psb_fasta_residues(const psb_fasta_t *fa) { return fa ? fa->cat.data() : nullptr; }
const int64_t *psb_fasta_offsets(const psb_fasta_t *fa) { return fa ? fa->off.data() : nullptr; }
// name of record i (not NUL-terminated): pointer and length
const char *psb_fasta_name(const psb_fasta_t *fa, int64_t i, int *len) {
    if (!fa || i < 0 || i >= fa->n) { if (len) *len = 0; return nullptr; }
    if (len) *len = (int)(fa->name_off[(size_t)i + 1] - fa->name_off[(size_t)i]);
    return fa->names.data() + fa->name_off[(size_t)i];
}
void psb_fasta_free(psb_fasta_t *fa) { delete fa; }

psb_db_t *psb_db_from_fasta(const char *path, const parasail_matrix_t *matrix) {
    psb_fasta_t *fa = psb_fasta_read(path);
    if (!fa) return nullptr;
    for (int64_t i = 0; i < fa->n; ++i)
        if (fa->off[(size_t)i + 1] == fa->off[(size_t)i]) {
            set_error("psb_db_from_fasta: record " + std::to_string(i) + " has no residues");
            psb_fasta_free(fa);
            return nullptr;
        }
    psb_db_t *db = psb_db_create(fa->cat.data(), fa->off.data(), fa->n, matrix);
    psb_fasta_free(fa);
    return db;
}

static const char kMagic[8] = {'P', 'S', 'B', 'D', 'B', 0, 1, 0};

int psb_db_save(const psb_db_t *db, const char *path) {
    if (!db || !path) { set_error("psb_db_save: NULL argument"); return PSB_EINVAL; }
    int dev = 0; cudaStream_t st = nullptr;
    const int rc = db_io_ensure_ctx(&dev, &st);
    if (rc != PSB_OK) return rc;
    if (dev != db->device) { set_error("psb_db_save: database lives on another device"); return PSB_EINVAL; }
    const size_t n = (size_t)db->n, nw = (size_t)db->words;
    std::vector<int> len(n), perm(n);
    std::vector<long long> woff(n + 1);
    std::vector<unsigned> words(nw);
    cudaError_t e = cudaStreamSynchronize(db->stream);
    if (e == cudaSuccess) e = cudaMemcpy(len.data(), db->d_len, n * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(perm.data(), db->d_perm, n * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(woff.data(), db->d_word_off, (n + 1) * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(words.data(), db->d_words, nw * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error(std::string("psb_db_save: ") + cudaGetErrorString(e)); return PSB_ECUDA; }
    std::FILE *f = std::fopen(path, "wb");
    if (!f) { set_error(std::string("psb_db_save: cannot create ") + path); return PSB_EINVAL; }
    const int32_t bits = db->bits, msize = db->msize;
    const int64_t hdr[3] = {db->n, db->residues, db->words}, reserved = 0;
    bool ok = std::fwrite(kMagic, 1, 8, f) == 8 && std::fwrite(&bits, 4, 1, f) == 1 && std::fwrite(&msize, 4, 1, f) == 1 &&
              std::fwrite(hdr, 8, 3, f) == 3 && std::fwrite(db->mapper, 1, 256, f) == 256 && std::fwrite(&reserved, 8, 1, f) == 1;
    ok = ok && std::fwrite(len.data(), 4, n, f) == n && std::fwrite(perm.data(), 4, n, f) == n &&
         std::fwrite(woff.data(), 8, n + 1, f) == n + 1 && std::fwrite(words.data(), 4, nw, f) == nw;
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) { set_error(std::string("psb_db_save: short write to ") + path); return PSB_EINVAL; }
    return PSB_OK;
}

psb_db_t *psb_db_load(const char *path, const parasail_matrix_t *matrix) {
    if (!path) { set_error("psb_db_load: NULL path"); return nullptr; }
    int dev = 0; cudaStream_t st = nullptr;
    if (db_io_ensure_ctx(&dev, &st) != PSB_OK) return nullptr;
    std::FILE *f = std::fopen(path, "rb");
    if (!f) { set_error(std::string("psb_db_load: cannot open ") + path); return nullptr; }
    char magic[8]; int32_t bits = 0, msize = 0; int64_t hdr[3] = {0, 0, 0}, reserved = 0;
    uint8_t mapper[256];
    bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, kMagic, 8) == 0 && std::fread(&bits, 4, 1, f) == 1 &&
              std::fread(&msize, 4, 1, f) == 1 && std::fread(hdr, 8, 3, f) == 3 && std::fread(mapper, 1, 256, f) == 256 &&
              std::fread(&reserved, 8, 1, f) == 1;
    if (!ok || (bits != 2 && bits != 3 && bits != 5) || hdr[0] <= 0 || hdr[0] > 0x7ffffffe || hdr[2] <= 0) {
        std::fclose(f);
        set_error(std::string("psb_db_load: not a packed database file: ") + path);
        return nullptr;
    }
    if (matrix) {
        HostMatrix hm(matrix);
        if (hm.size != msize || std::memcmp(hm.mapper, mapper, 256) != 0) {
            std::fclose(f);
            set_error("psb_db_load: the file was packed with a different alphabet than this matrix");
            return nullptr;
        }
    }
    const size_t n = (size_t)hdr[0], nw = (size_t)hdr[2];
    std::vector<int> len(n), perm(n);
    std::vector<long long> woff(n + 1);
    std::vector<unsigned> words(nw);
    ok = std::fread(len.data(), 4, n, f) == n && std::fread(perm.data(), 4, n, f) == n &&
         std::fread(woff.data(), 8, n + 1, f) == n + 1 && std::fread(words.data(), 4, nw, f) == nw;
    std::fclose(f);
    const int rpw = db_residues_per_word(bits);
    for (size_t i = 0; ok && i < n; ++i)
        ok = len[i] > 0 && (i == 0 || len[i] <= len[i - 1]) && perm[i] >= 0 && (size_t)perm[i] < n &&
             woff[i + 1] - woff[i] == (len[i] + rpw - 1) / rpw;
    ok = ok && woff[0] == 0 && (size_t)woff[n] == nw;
    if (!ok) { set_error(std::string("psb_db_load: truncated or inconsistent file: ") + path); return nullptr; }
    psb_db *db = new psb_db();
    db->device = dev; db->stream = st; db->n = (int64_t)n; db->residues = hdr[1]; db->words = (int64_t)nw; db->bits = bits; db->msize = msize;
    std::memcpy(db->mapper, mapper, 256);
    db->maxlen = len[0];
    db->host_len.assign(n, 0);
    for (size_t i = 0; i < n; ++i) { db->host_len[(size_t)perm[i]] = len[i]; if (len[i] > 65535) db->nlong++; }
    db->top_len.assign(len.begin(), len.begin() + std::min<size_t>(n, 4096));
    cudaError_t e = cudaMallocAsync(&db->d_word_off, (n + 1) * 8, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&db->d_perm, n * 4, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&db->d_len, n * 4, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&db->d_words, nw * 4 + 64, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db->d_word_off, woff.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db->d_perm, perm.data(), n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db->d_len, len.data(), n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db->d_words, words.data(), nw * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // the host vectors go out of scope
    if (e != cudaSuccess) { set_error(std::string("psb_db_load: ") + cudaGetErrorString(e)); psb_db_free(db); return nullptr; }
    return db;
}

}  // extern "C"
