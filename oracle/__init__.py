"""CPU oracle for the colour hot path of fegemo/palette-and-histo-gan.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`palette_and_histo_gan_b200/`) may import this.  The only legitimate callers are
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs.

PARITY UNPINNED: the reference (TensorFlow 2.9.1, `requirements.txt:99`) cannot be imported in
this image (no TF wheel, no network) and ships neither tests nor golden vectors, so this oracle
is a restatement checked only against (a) the survey-time known-answer values recorded in
SURVEY.md §4, (b) an independent torch-autograd restatement (`oracle/torch_port.py`), (c) central
finite differences for the analytic gradient, and (d) the self-consistency identities the
reference's own call sites rely on (round trips, sum-to-one, symmetry), and — the strongest pin —
(e) the outputs of the reference's OWN source files (histogram.py, io_utils.py, dataset_utils.py, imported
unmodified) executed over a minimal TensorFlow-op shim (`oracle/ref_shim.py`, `oracle/run_reference.py`,
fixture `tests/golden/reference_run.npz`).  That pins the restatement to the reference's code; the semantics of
the individual TensorFlow ops remain an assumption until a run of real TensorFlow is possible.
"""
