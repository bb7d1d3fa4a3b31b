"""CPU oracle for the KDE score/draw hot path of vargatn/skysampler.

TEST INFRASTRUCTURE ONLY.  Nothing under ``skysampler_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` do, and there only as the checker (or as the timed CPU arm), never as the product.

What it restates, and what pins it
----------------------------------
The reference (``/root/reference``, pure Python) delegates every FLOP of this path to
scikit-learn, a third-party dependency that is **not vendored, not pinned and not even listed**
by the reference (``setup.py:11``).  The version present in this image is scikit-learn 1.9.0.

* ``oracle.brute64``   – fp64 brute-force restatement of the published definition
  ``log p(z) = log_knorm - log(sum w) + log sum_j w_j exp(-|z-z_j|^2 / 2h^2)``
  (sklearn ``neighbors/_binary_tree.pxi.tp:376-378,448-478,1527-1529``; ``_kde.py:273-287``).
* ``oracle.container`` – restatement of ``skysampler/emulator.py:234-394`` (``KDEContainer``):
  whitening, Jacobian, draw loop with radial window, ``score_samples`` — on top of sklearn's
  ``KernelDensity`` exactly as the reference calls it (``emulator.py:354-356,360,372,384``).
  This is the "port" that ``bench.py`` times on the host cores.
* ``oracle.draw``      – restatement of ``KernelDensity.sample`` (sklearn ``_kde.py:338-349``)
  from explicit uniform / normal streams.
* ``oracle.accept``    – restatement of the accept test of ``read_concentric``
  (``emulator.py:737-746``).

Parity status: **the reference ships no tests, golden vectors or fixtures for this path**
(SURVEY.md §4, §8c), so on its own the oracle would be "parity unpinned".  It is pinned instead
against outputs of the reference itself run in the build container: ``oracle/make_golden.py``
imports ``/root/reference/skysampler/emulator.py`` (with ``fitsio``/``matplotlib`` stubbed — both
unused on this path) and writes ``tests/golden/*.npz``; ``tests/test_oracle.py`` checks every
function here against those files.  ``/root/reference`` is never read at test/bench run time.
"""
