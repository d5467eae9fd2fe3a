"""CPU oracle for the FeTA spectral hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch/NumPy restatement, on the CPU, of the reference
algorithm (ansonb/FeTA_TMLR) for the path SURVEY.md section 8 names.  Every
function cites the reference file:line it follows.  It is the *checker* for the
CUDA product path in ``feta_tmlr_b200``; nothing in the product may import it.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import from here.

PARITY STATUS: **parity unpinned** for every floating-point function.  The
reference ships no tests, no golden vectors and no fixtures, it
cannot be imported in this image (torch_geometric / torch_scatter / ogb are
absent) and its attention layer source is missing from the tree
(SURVEY.md F1).  The third-party arithmetic the path rests on is
torch_geometric==1.7 (README.md:24 of the reference; ``get_laplacian``,
``remove_self_loops``, ``add_self_loops``, ``add_remaining_self_loops``,
``gcn_norm``, ``GCNConv``, ``global_mean_pool``, ``MessagePassing.propagate``)
whose published algorithm is restated in ``oracle/pyg17.py``.  The substitute
pinning is (a) an independent dense-matrix cross-check (``oracle/dense.py``),
(b) closed-form known-answer cases, (c) ``torch.autograd.gradcheck`` in fp64 and
(d) frozen golden vectors under ``tests/golden`` generated *from this oracle*
by ``tests/golden/make_golden.py``.  The integer paths (collate / index
builders) are fully determined by ``transformer/data.py`` and restated exactly.
"""
