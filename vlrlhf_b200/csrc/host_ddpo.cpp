// Host-side (CPU, integer) DDPO token mask -- the `mask_shared_tokens` branch of VLDPOTrainer.get_batch_logps
// (base/trainer.py:169-184) over utils/diff_lib.get_diff_ids (utils/diff_lib.py:116-180).
//
// The reference calls Python's difflib.SequenceMatcher(None, a, b).get_matching_blocks() (autojunk ON, no junk
// predicate) once per pair inside the step.  difflib is CPython standard library, not under /root/reference; its
// algorithm is restated here exactly (Lib/difflib.py __chain_b / find_longest_match / get_matching_blocks):
//   * b2j: element -> ascending positions in b; when len(b) >= 200 every element occurring more than
//     len(b)//100 + 1 times is "popular" and dropped from b2j (it can no longer SEED a match, but the extension
//     loops still run over it because the junk set is empty);
//   * find_longest_match: row-by-row longest-common-substring DP over b2j hits, strict `>` so the first best match in
//     (i ascending, j ascending) order wins, then extension to both sides while elements are equal;
//   * get_matching_blocks: LIFO work list, sort, merge adjacent blocks, (la, lb, 0) sentinel.
// On top: keep blocks >= min_match_size (+ sentinel), the gaps between kept blocks that are non-empty on BOTH
// sides are the modified spans (diff_lib.py:136-163), and the spans are mapped from the merged shifted label layout
// back to text-level logits rows (row j-1 predicts text token j) as vlrlhf_b200/host.py:ddpo_row_weights does.
//
// This is a HOST function of the C ABI (plain host pointers); it needs no GPU and launches nothing.
#include <stdint.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <unordered_map>
#include <vector>

#include "../../include/vlb200.h"

namespace vlb {
void set_last_error(const char* fmt, ...);
}

namespace {

struct Block { int i, j, k; };

struct Matcher {
    const std::vector<int64_t>& a;
    const std::vector<int64_t>& b;
    std::unordered_map<int64_t, std::vector<int>> b2j;
    std::vector<int> cur, prev, touched_cur, touched_prev;

    Matcher(const std::vector<int64_t>& a_, const std::vector<int64_t>& b_) : a(a_), b(b_) {
        const int n = (int)b.size();
        b2j.reserve(n * 2 + 1);
        for (int i = 0; i < n; ++i) b2j[b[i]].push_back(i);
        if (n >= 200) {  // autojunk
            const size_t ntest = (size_t)(n / 100 + 1);
            for (auto it = b2j.begin(); it != b2j.end();) {
                if (it->second.size() > ntest) it = b2j.erase(it);
                else ++it;
            }
        }
        cur.assign(n + 1, 0);
        prev.assign(n + 1, 0);
    }

    Block find_longest_match(int alo, int ahi, int blo, int bhi) {
        int besti = alo, bestj = blo, bestsize = 0;
        // j2len[j] lives at index j+1 so that j-1 == -1 is addressable; only touched entries are non-zero
        for (int t : touched_prev) prev[t] = 0;
        touched_prev.clear();
        for (int i = alo; i < ahi; ++i) {
            touched_cur.clear();
            auto it = b2j.find(a[i]);
            if (it != b2j.end()) {
                for (int j : it->second) {
                    if (j < blo) continue;
                    if (j >= bhi) break;
                    const int k = prev[j] + 1;  // prev[(j-1)+1]
                    cur[j + 1] = k;
                    touched_cur.push_back(j + 1);
                    if (k > bestsize) { besti = i - k + 1; bestj = j - k + 1; bestsize = k; }
                }
            }
            for (int t : touched_prev) prev[t] = 0;
            std::swap(cur, prev);
            std::swap(touched_cur, touched_prev);
        }
        for (int t : touched_prev) prev[t] = 0;
        touched_prev.clear();
        while (besti > alo && bestj > blo && a[besti - 1] == b[bestj - 1]) { --besti; --bestj; ++bestsize; }
        while (besti + bestsize < ahi && bestj + bestsize < bhi && a[besti + bestsize] == b[bestj + bestsize]) ++bestsize;
        return {besti, bestj, bestsize};
    }

    std::vector<Block> matching_blocks() {
        const int la = (int)a.size(), lb = (int)b.size();
        struct Range { int alo, ahi, blo, bhi; };
        std::vector<Range> queue{{0, la, 0, lb}};
        std::vector<Block> blocks;
        while (!queue.empty()) {
            const Range r = queue.back();
            queue.pop_back();
            const Block x = find_longest_match(r.alo, r.ahi, r.blo, r.bhi);
            if (x.k) {
                blocks.push_back(x);
                if (r.alo < x.i && r.blo < x.j) queue.push_back({r.alo, x.i, r.blo, x.j});
                if (x.i + x.k < r.ahi && x.j + x.k < r.bhi) queue.push_back({x.i + x.k, r.ahi, x.j + x.k, r.bhi});
            }
        }
        std::sort(blocks.begin(), blocks.end(), [](const Block& p, const Block& q) {
            return p.i != q.i ? p.i < q.i : (p.j != q.j ? p.j < q.j : p.k < q.k);
        });
        std::vector<Block> out;
        int i1 = 0, j1 = 0, k1 = 0;
        for (const Block& x : blocks) {
            if (i1 + k1 == x.i && j1 + k1 == x.j) {
                k1 += x.k;
            } else {
                if (k1) out.push_back({i1, j1, k1});
                i1 = x.i; j1 = x.j; k1 = x.k;
            }
        }
        if (k1) out.push_back({i1, j1, k1});
        out.push_back({la, lb, 0});
        return out;
    }
};

// merged shifted label sequence of one row + for every text token j the index of its label in that sequence (-1: none)
void merged_shift(const int64_t* ids, const int64_t* am, const int64_t* lab, int L, int image_token, int feat_len,
                  int merged_len, int64_t label_pad, std::vector<int64_t>& seq, std::vector<int>& row_pos) {
    seq.clear();
    row_pos.assign(L, -1);
    for (int j = 0; j < L; ++j) {
        if (am && am[j] == 0) continue;
        if (ids[j] == image_token) {
            row_pos[j] = (int)seq.size() - 1;
            seq.insert(seq.end(), (size_t)feat_len, 0);
        } else {
            row_pos[j] = (int)seq.size() - 1;
            seq.push_back(lab[j] == label_pad ? 0 : lab[j]);
        }
    }
    if (merged_len >= 0 && (int)seq.size() < merged_len) seq.insert(seq.end(), (size_t)(merged_len - (int)seq.size()), 0);
    if (!seq.empty()) seq.erase(seq.begin());  // labels[:, 1:]
}

}  // namespace

extern "C" int vlb200_host_matching_blocks(const int64_t* a, int na, const int64_t* b, int nb, int* triples, int max_blocks) {
    if ((!a && na) || (!b && nb) || !triples || na < 0 || nb < 0) {
        vlb::set_last_error("host_matching_blocks: bad arguments");
        return -1;
    }
    std::vector<int64_t> va(a, a + na), vb(b, b + nb);
    Matcher m(va, vb);
    const std::vector<Block> out = m.matching_blocks();
    if ((int)out.size() > max_blocks) {
        vlb::set_last_error("host_matching_blocks: %d blocks do not fit max_blocks=%d", (int)out.size(), max_blocks);
        return -1;
    }
    for (size_t t = 0; t < out.size(); ++t) { triples[3 * t] = out[t].i; triples[3 * t + 1] = out[t].j; triples[3 * t + 2] = out[t].k; }
    return (int)out.size();
}

extern "C" int vlb200_host_ddpo_row_weights(const int64_t* input_ids, const int64_t* attention_mask, const int64_t* labels,
                                            int n_seq, int text_len, int image_token, const int* feat_len_per_seq,
                                            int merged_len, int64_t label_pad_token_id, int min_match_size,
                                            uint8_t* weights) {
    if (!input_ids || !labels || !feat_len_per_seq || !weights || n_seq <= 0 || n_seq % 2 || text_len < 2) {
        vlb::set_last_error("host_ddpo_row_weights: bad arguments (n_seq must be even: chosen rows then rejected rows)");
        return VLB200_ERR_INVALID;
    }
    const int n = n_seq / 2, L = text_len;
    std::fill(weights, weights + (size_t)n_seq * (L - 1), (uint8_t)0);
    std::vector<int64_t> sa, sb;
    std::vector<int> pa, pb;
    std::vector<uint8_t> ma, mb;
    for (int p = 0; p < n; ++p) {
        const int ra = p, rb = n + p;
        merged_shift(input_ids + (size_t)ra * L, attention_mask ? attention_mask + (size_t)ra * L : nullptr,
                     labels + (size_t)ra * L, L, image_token, feat_len_per_seq[ra], merged_len, label_pad_token_id, sa, pa);
        merged_shift(input_ids + (size_t)rb * L, attention_mask ? attention_mask + (size_t)rb * L : nullptr,
                     labels + (size_t)rb * L, L, image_token, feat_len_per_seq[rb], merged_len, label_pad_token_id, sb, pb);
        Matcher m(sa, sb);
        std::vector<Block> blocks = m.matching_blocks();
        ma.assign(sa.size(), 0);
        mb.assign(sb.size(), 0);
        int ai = 0, bi = 0;
        for (size_t t = 0; t < blocks.size(); ++t) {
            const Block& x = blocks[t];
            if (t + 1 < blocks.size() && x.k < min_match_size) continue;  // short matches count as modified
            if (x.i > ai && x.j > bi) {  // the gap before this block is non-empty on both sides
                std::fill(ma.begin() + ai, ma.begin() + x.i, (uint8_t)1);
                std::fill(mb.begin() + bi, mb.begin() + x.j, (uint8_t)1);
            }
            ai = x.i + x.k;
            bi = x.j + x.k;
        }
        for (int j = 1; j < L; ++j) {
            if (pa[j] >= 0 && pa[j] < (int)ma.size() && ma[pa[j]]) weights[(size_t)ra * (L - 1) + j - 1] = 1;
            if (pb[j] >= 0 && pb[j] < (int)mb.size() && mb[pb[j]]) weights[(size_t)rb * (L - 1) + j - 1] = 1;
        }
    }
    return VLB200_OK;
}
