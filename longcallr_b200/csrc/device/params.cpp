/* params.cpp — preset defaults, src/main.rs:272-396 (values listed in SURVEY.md section 5). */
#include <string.h>

#include "longcallr_b200.h"

extern "C" int lcr_params_preset(int preset, lcr_params *p) {
    if (!p || preset < 0 || preset > 3) return LCR_ERR_INVALID_ARG;
    memset(p, 0, sizeof *p);
    const bool ont = preset == LCR_PRESET_ONT_CDNA || preset == LCR_PRESET_ONT_DRNA;
    p->platform = ont ? 1 : 0;                       /* main.rs:274,305,336,367 */
    p->min_depth = ont ? 10 : 6;                     /* :275,306,337,368 */
    p->min_phase_score = ont ? 13.0f : 11.0f;        /* :276,307,338,369 */
    p->read_assignment_cutoff = 0.0;                 /* min_read_assignment_diff */
    p->min_linkers = 1;
    p->min_allele_freq = ont ? 0.20f : 0.15f;        /* :279,310,341,372 */
    p->min_allele_freq_include_intron = 0.0f;
    p->distance_to_read_end = ont ? 20 : 40;         /* :282,313,344,375 */
    p->dense_win_size = 100;
    p->min_dense_cnt = 5;
    p->use_strand_bias = (preset == LCR_PRESET_ONT_CDNA || preset == LCR_PRESET_HIFI_ISOSEQ) ? 1 : 0; /* :285,316,347,378 */
    p->max_enum_snps = 10;
    p->min_mapq = 20;
    p->divergence = 0.5f;
    p->min_baseq = 10;
    p->min_qual = 2;
    p->polya_tail_length = 5;
    p->max_depth = 50000;
    p->min_read_length = 500;
    p->low_allele_frac_cutoff = 0.05f;
    p->low_allele_cnt_cutoff = 10;
    p->ld_weight_threshold = 1;                      /* thread.rs:166 */
    p->flags = 0;
    p->seed = 0;
    p->downsample_depth = 10000;                     /* main.rs:296,327,358,389 */
    return LCR_OK;
}
