// gpba_order.h -- host-only symbolic phase of the reduced-system factorization (no CUDA): fill-reducing order of the
// pose blocks, tile-level symbolic factorization and level schedule.  It is what SimplicialLDLT::analyzePattern + AMD do
// for the reference (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:147-201), restated for 48 x 48 tiles.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

namespace gpba {

struct CholSymbolic {
  int NT = 0;                         // tile columns (padding positions included)
  int n_parts = 1;                    // leaves + separators of the nested-dissection order
  int n_levels = 0;                   // length of the dependency chain of tile columns
  int64_t doubles = 0;                // storage of the non-zero tiles
  std::vector<int> perm;              // [n_pose] position (pose-block slot, gaps included) of pose block b
  std::vector<unsigned char> pos_used;  // [NT * bpt]
  std::vector<int> col_begin, col_rows; // per tile column: rows i > k with L_ik != 0
  std::vector<int> row_begin, row_cols; // the same pattern by rows
  std::vector<int> level;             // [NT] level of a tile column: it depends on column j iff tile (k, j) != 0
  std::vector<int64_t> tile_off;      // [NT * NT] offset of tile (i, j), i >= j, or -1
};

// hs_row / hs_col: upper block pattern (row <= col) of the reduced system; tile = bpt pose blocks.
inline void chol_symbolic(int n_pose, int n_hs, const int* hs_row, const int* hs_col, int bpt, int tile_doubles, int max_depth, bool verbose, CholSymbolic& out) {
  int& NT = out.NT;
  int chol_parts = 1;
  // Fill-reducing order of the pose blocks (the role AMD plays for SimplicialLDLT, linear_solver_eigen.h:147-201).
  // Small systems (a local window): reverse Cuthill-McKee on the Hschur block graph, so that a revisited place sits next
  // to its first visit and the factor stays banded.  Large systems: nested dissection by BFS level structures -- the level
  // set nearest to the middle of the level structure rooted at a pseudo-peripheral node separates what comes before it
  // from what comes after it (George's automatic nested dissection); the two sides are ordered first (recursively), the
  // separator last.  The sides are independent, so their tile columns land in the same levels of the schedule below and
  // run in the same launches: the factorization is bound by the LENGTH of the dependency chain (one tile column = one
  // 48-pivot chain plus two kernel boundaries), and the chain is now the deepest leaf plus the separators above it
  // instead of all columns (C4: 250 -> 70 levels for 1.4x the flops; offline study in profiles/r02_ordering_study.txt).
  // Every part (leaf or separator) starts on a tile boundary, otherwise a tile shared by two leaves would chain them
  // together; the positions skipped for that are padded with identity rows like the tail of the matrix.
  std::vector<int> perm(n_pose, 0);   // permuted POSITION (in pose blocks, gaps included) of pose block b
  int n_positions = 0;
  {
    std::vector<std::vector<int>> adj(n_pose);
    for (int k = 0; k < n_hs; ++k) if (hs_row[k] != hs_col[k]) { adj[hs_row[k]].push_back(hs_col[k]); adj[hs_col[k]].push_back(hs_row[k]); }
    std::vector<int> mark(n_pose, -1), lev(n_pose, 0);   // mark[v] == id: v belongs to the node set `id` being processed
    int next_id = 0;
    // BFS over the nodes marked `id`: fills lev[] and returns the visiting order (stamp[] = visited in this call)
    std::vector<int> stamp(n_pose, -1);
    int stamp_id = 0;
    auto bfs_levels = [&](int id, int start, std::vector<int>& order) {
      ++stamp_id;
      order.clear();
      order.push_back(start); lev[start] = 0; stamp[start] = stamp_id;
      for (size_t head = 0; head < order.size(); ++head) {
        const int v = order[head];
        for (int w : adj[v]) if (mark[w] == id && stamp[w] != stamp_id) { stamp[w] = stamp_id; lev[w] = lev[v] + 1; order.push_back(w); }
      }
    };
    // pseudo-peripheral node of the connected component of `start`: restart from the farthest node while the depth grows
    auto peripheral = [&](int id, int start, std::vector<int>& order) {
      bfs_levels(id, start, order);
      for (int it = 0; it < 4; ++it) {
        const int depth = lev[order.back()], far = order.back();
        std::vector<int> o2;
        bfs_levels(id, far, o2);
        const bool deeper = lev[o2.back()] > depth;
        order.swap(o2);
        if (!deeper) break;
      }
    };
    // reverse Cuthill-McKee of the nodes marked `id` (all components), appended to out
    auto rcm = [&](int id, const std::vector<int>& nodes, std::vector<int>& out) {
      std::vector<int> order, comp;
      std::vector<char> ordered(n_pose, 0);
      for (int s0 : nodes) {
        if (ordered[s0]) continue;
        // component of s0 among the not yet ordered nodes: temporarily re-mark it
        const int cid = next_id++;
        comp.clear(); comp.push_back(s0); mark[s0] = cid;
        for (size_t head = 0; head < comp.size(); ++head)
          for (int w : adj[comp[head]]) if (mark[w] == id && !ordered[w]) { mark[w] = cid; comp.push_back(w); }
        peripheral(cid, s0, order);
        const int root = order[0];
        // Cuthill-McKee from the root: neighbours by ascending degree
        std::vector<int> cm; cm.push_back(root);
        ++stamp_id; stamp[root] = stamp_id;
        for (size_t head = 0; head < cm.size(); ++head) {
          const int v = cm[head];
          std::vector<int> nb;
          for (int w : adj[v]) if (mark[w] == cid && stamp[w] != stamp_id) { stamp[w] = stamp_id; nb.push_back(w); }
          std::sort(nb.begin(), nb.end(), [&](int x, int y) { return adj[x].size() != adj[y].size() ? adj[x].size() < adj[y].size() : x < y; });
          cm.insert(cm.end(), nb.begin(), nb.end());
        }
        for (size_t i = cm.size(); i-- > 0;) { out.push_back(cm[i]); ordered[cm[i]] = 1; }
        for (int v : comp) mark[v] = id;
      }
    };
    std::vector<std::vector<int>> parts;   // elimination order: parts in sequence, nodes inside a part in sequence
    if (max_depth < 0) max_depth = n_pose >= 256 ? 8 : 0;   // default: local windows keep the plain banded order
    std::function<void(std::vector<int>&, int)> dissect = [&](std::vector<int>& nodes, int depth) {
      const int id = next_id++;
      for (int v : nodes) mark[v] = id;
      auto leaf = [&]() { parts.emplace_back(); rcm(id, nodes, parts.back()); };
      if (depth <= 0 || (int)nodes.size() < 16 * bpt) { leaf(); return; }
      // connected components are independent parts
      std::vector<int> order;
      bfs_levels(id, nodes[0], order);
      if (order.size() < nodes.size()) {
        std::vector<std::vector<int>> comps;
        std::vector<char> got(n_pose, 0);
        for (int s0 : nodes) {
          if (got[s0]) continue;
          bfs_levels(id, s0, order);
          for (int v : order) got[v] = 1;
          comps.push_back(order);
        }
        for (auto& c : comps) dissect(c, depth);
        return;
      }
      peripheral(id, nodes[0], order);
      const int L = lev[order.back()];
      if (L < 4) { leaf(); return; }
      std::vector<int> size(L + 1, 0);
      for (int v : order) size[lev[v]]++;
      // level whose middle is nearest to half of the nodes
      int best = 1; double best_d = 1e300; int64_t cum = 0;
      for (int l = 0; l <= L; ++l) {
        const double mid = (double)cum + 0.5 * size[l];
        cum += size[l];
        if (l >= 1 && l < L) { const double d2 = std::fabs(mid - 0.5 * (double)nodes.size()); if (d2 < best_d) { best_d = d2; best = l; } }
      }
      std::vector<int> A, B, S;
      for (int v : order) {
        if (lev[v] < best) A.push_back(v);
        else if (lev[v] > best) B.push_back(v);
        else {
          bool touches_b = false;
          for (int w : adj[v]) if (mark[w] == id && lev[w] > best) { touches_b = true; break; }
          (touches_b ? S : A).push_back(v);   // a node of the level that no later node touches separates nothing
        }
      }
      // worthwhile only while the sides stay clearly larger than what separates them (the separator's rows are carried
      // through every column of its sides as fill)
      if ((int)nodes.size() < 4 * (int)S.size() || A.empty() || B.empty()) { leaf(); return; }
      dissect(A, depth - 1);
      dissect(B, depth - 1);
      const int sid = next_id++;
      for (int v : S) mark[v] = sid;
      parts.emplace_back();
      rcm(sid, S, parts.back());
    };
    std::vector<int> all(n_pose);
    for (int i = 0; i < n_pose; ++i) all[i] = i;
    if (n_pose > 0) dissect(all, max_depth);
    int cursor = 0;
    for (auto& part : parts) {
      for (int v : part) perm[v] = cursor++;
      cursor = (cursor + bpt - 1) / bpt * bpt;   // the next part starts a new tile
    }
    n_positions = cursor;
    chol_parts = (int)parts.size();
    if (verbose) fprintf(stderr, "[gpba] cholesky: nested dissection depth <= %d: %d parts, %d positions for %d pose blocks\n", max_depth, chol_parts, n_positions, n_pose);
  }

  NT = std::max(1, n_positions / bpt);
  out.n_parts = chol_parts;
  out.perm = perm;
  out.pos_used.assign((size_t)NT * bpt, 0);
  for (int i = 0; i < n_pose; ++i) out.pos_used[perm[i]] = 1;
  // tile-level symbolic factorization: right-looking fill of the lower triangle
  std::vector<char> nz((size_t)NT * NT, 0);
  for (int t = 0; t < NT; ++t) nz[(size_t)t * NT + t] = 1;
  for (int k = 0; k < n_hs; ++k) {
    int ti = perm[hs_col[k]] / bpt, tj = perm[hs_row[k]] / bpt;
    if (ti < tj) std::swap(ti, tj);  // lower triangle
    nz[(size_t)ti * NT + tj] = 1;
  }
  out.col_rows.clear();
  out.col_begin.assign(NT + 1, 0);
  std::vector<int> rows;
  for (int k = 0; k < NT; ++k) {
    rows.clear();
    for (int i = k + 1; i < NT; ++i) if (nz[(size_t)i * NT + k]) rows.push_back(i);
    for (size_t a = 0; a < rows.size(); ++a)
      for (size_t b = 0; b <= a; ++b) nz[(size_t)rows[a] * NT + rows[b]] = 1;  // fill
    out.col_begin[k] = (int)out.col_rows.size();
    out.col_rows.insert(out.col_rows.end(), rows.begin(), rows.end());
  }
  out.col_begin[NT] = (int)out.col_rows.size();
  out.tile_off.assign((size_t)NT * NT, -1);
  int64_t cursor2 = 0;
  for (int i = 0; i < NT; ++i)
    for (int j = 0; j <= i; ++j)
      if (nz[(size_t)i * NT + j]) { out.tile_off[(size_t)i * NT + j] = cursor2; cursor2 += tile_doubles; }
  out.doubles = cursor2;
  // row-wise view of the factor pattern for the backward substitution
  out.row_begin.assign(NT + 1, 0);
  for (int r : out.col_rows) out.row_begin[r + 1]++;
  for (int i = 0; i < NT; ++i) out.row_begin[i + 1] += out.row_begin[i];
  out.row_cols.assign(out.col_rows.size(), 0);
  std::vector<int> cur_r(out.row_begin.begin(), out.row_begin.end() - 1);
  for (int k = 0; k < NT; ++k)
    for (int e = out.col_begin[k]; e < out.col_begin[k + 1]; ++e) out.row_cols[cur_r[out.col_rows[e]]++] = k;
  // level schedule: column r depends on column j < r iff tile (r, j) is non-zero; columns of equal level are independent
  out.level.assign(NT, 0);
  out.n_levels = 0;
  for (int j = 0; j < NT; ++j) {
    for (int e = out.col_begin[j]; e < out.col_begin[j + 1]; ++e) out.level[out.col_rows[e]] = std::max(out.level[out.col_rows[e]], out.level[j] + 1);
    out.n_levels = std::max(out.n_levels, out.level[j] + 1);
  }
}

}  // namespace gpba
