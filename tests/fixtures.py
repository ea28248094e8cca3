"""Small hand-made inverted index shared by the oracle tests and the GPU parity tests."""
import numpy as np


def tiny_index():
    """docs 0..5, terms 0..3; body and title postings with positions (incl. the -100 sentinel)."""
    body = {0: {0: (0.5, [1, 7]), 1: (1.0, [2]), 4: (0.25, [9])},
            1: {0: (1.0, [2, 30]), 1: (0.5, [5]), 2: (1.0, [0])},
            2: {0: (0.75, [3]), 3: (1.0, [4])},
            3: {5: (1.0, [])}}
    title = {0: {1: (1.0, [-100]), 2: (0.5, [0])},
             1: {2: (1.0, [1]), 3: (1.0, [-100])},
             2: {0: (1.0, [-100]), 1: (0.5, [-100])}}

    def csc(tab, n_terms):
        ptr, docs, tf, pp, pos = [0], [], [], [0], []
        for t in range(n_terms):
            for d in sorted(tab.get(t, {})):
                docs.append(d)
                tf.append(tab[t][d][0])
                pos.extend(tab[t][d][1])
                pp.append(len(pos))
            ptr.append(len(docs))
        return (np.array(ptr, np.uint64), np.array(docs, np.uint32), np.array(tf, np.float32),
                np.array(pp, np.uint64), np.array(pos, np.float32))

    return csc(title, 4), csc(body, 4), 6


def queries_csr(kw_lists, ph_lists=None):
    kw_ptr = np.zeros(len(kw_lists) + 1, np.uint64)
    kw_ptr[1:] = np.cumsum([len(x) for x in kw_lists])
    kw = np.array([t for x in kw_lists for t in x], np.uint32)
    if ph_lists is None:
        return kw_ptr, kw, None, None
    ph_ptr = np.zeros(len(ph_lists) + 1, np.uint64)
    ph_ptr[1:] = np.cumsum([len(x) for x in ph_lists])
    ph = np.array([t for x in ph_lists for t in x], np.uint32)
    return kw_ptr, kw, ph_ptr, ph
