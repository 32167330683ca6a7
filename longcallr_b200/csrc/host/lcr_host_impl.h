/* lcr_host_impl.h — owning containers behind the views of lcr_host.h (internal). */
#ifndef LCR_HOST_IMPL_H
#define LCR_HOST_IMPL_H

#include <cstddef>
#include <string>
#include <vector>

#include "lcr_host.h"

namespace lcrhost {

struct Reads {
    lcr_reads view{}; /* handed out; the owner is recovered with offsetof */
    std::vector<std::string> contig_names;
    std::vector<const char *> name_ptrs;
    std::vector<uint64_t> contig_lens;
    std::vector<int32_t> tid, pos;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> mapq;
    std::vector<int8_t> ts;
    std::vector<float> de;
    std::vector<uint64_t> seq_off, cig_off, qname_off;
    std::vector<uint8_t> seq, qual;
    std::vector<uint32_t> cigar;
    std::vector<char> qnames;
    void finish();
};

struct Fasta {
    lcr_fasta view{};
    std::vector<std::string> names;
    std::vector<std::vector<uint8_t>> seqs;
    std::vector<const char *> name_ptrs;
    std::vector<const uint8_t *> seq_ptrs;
    std::vector<uint64_t> lens;
    void finish();
};

struct Regions {
    lcr_region_list view{};
    std::vector<lcr_region> regions;
    std::vector<uint32_t> max_coverage;
};

bool read_file(const char *path, std::vector<uint8_t> &buf);
bool bgzf_inflate(const std::vector<uint8_t> &file, std::vector<uint8_t> &out, int n_threads);
int read_bam(const char *path, int n_threads, Reads &R);
int read_fasta(const char *path, Fasta &F);
int find_regions(const lcr_reads &R, const lcr_params &P, bool truncation, uint32_t trunc_cov, Regions &out);

} // namespace lcrhost
#endif
