"""TEST INFRASTRUCTURE -- numpy restatement of the user-history graph construction.

Follows reference ``MIND_corpus.py:162-216`` (one behaviour line -> dense normalised adjacency,
category mask, per-slot cluster indices).  The loops are kept in the same order as the reference
on purpose: this file is the bit-exact checker for ``nnr_sue_graph_build``.
"""
import numpy as np


def build_history_graph(categories, max_history_num, category_num, no_self_connection=False,
                        no_adjacent_normalization=False, gcn_normalization_type='symmetric'):
    """categories: sequence of category ids of the (already truncated, <= max_history_num) history.

    Returns (graph [G,G] float32, category_mask [C+1] bool, category_indices [H] int64), with
    G = max_history_num + category_num.  Reference: MIND_corpus.py:178-213.
    """
    H, C = int(max_history_num), int(category_num)
    G = H + C
    cats = [int(c) for c in categories][:H]
    n = len(cats)
    if no_self_connection:                                   # MIND_corpus.py:179-182
        A = np.zeros([G, G], dtype=np.float32)
    else:
        A = np.identity(G, dtype=np.float32)
    mask = np.zeros(C + 1, dtype=bool)                       # :183 (extra slot = padding cluster)
    idx = np.full([H], C, dtype=np.int64)                    # :184
    if n > 0:                                                # :185
        for i in range(n):
            ci = cats[i]
            mask[ci] = True                                  # :191
            idx[i] = ci                                      # :192
            A[i, H + ci] = 1                                 # :193-194 news <-> its proxy node
            A[H + ci, i] = 1
            for j in range(i + 1, n):
                cj = cats[j]
                if ci == cj:                                 # :197-199 intra-cluster clique
                    A[i, j] = 1
                    A[j, i] = 1
                else:                                        # :200-202 proxy <-> proxy
                    A[H + ci, H + cj] = 1
                    A[H + cj, H + ci] = 1
        if not no_adjacent_normalization:                    # :203
            if gcn_normalization_type == 'asymmetric':       # :204-208  D^-1 A
                D_inv = np.zeros([G, G], dtype=np.float32)
                np.fill_diagonal(D_inv, 1 / A.sum(axis=1, keepdims=False))
                A = np.matmul(D_inv, A)
            else:                                            # :209-213  D^-1/2 A D^-1/2
                D = np.zeros([G, G], dtype=np.float32)
                np.fill_diagonal(D, np.sqrt(1 / A.sum(axis=1, keepdims=False)))
                A = np.matmul(np.matmul(D, A), D)
    return A, mask, idx


def build_batch(category_lists, max_history_num, category_num, **kw):
    g, m, i = zip(*[build_history_graph(c, max_history_num, category_num, **kw) for c in category_lists])
    return np.stack(g), np.stack(m), np.stack(i)
