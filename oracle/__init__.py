"""CPU oracle for the sos-b200 hot path.  TEST INFRASTRUCTURE ONLY.

Everything in this package is a plain numpy / torch-CPU restatement of the
reference algorithm (henryxrl/Listening-to-Sound-of-Silence-for-Speech-Denoising)
for the one hot path this repo accelerates.  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` -- never by the product package.  The product path has no CPU
fallback: it raises if the CUDA library is missing.

Pinning status: the reference ships no tests and no golden vectors for this
path, so the restatement is pinned against the
reference ITSELF: ``oracle/make_golden.py`` imports the reference modules from
``/root/reference`` (networks.py of both models, which only need torch, and
transform.py with a stub ``librosa`` module whose stft/istft are this oracle's
numpy restatement of librosa 0.7.1) and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every oracle function against them.
librosa==0.7.1 (requirements.txt:4) itself is a third-party dependency that is
absent from /root/reference and cannot be installed offline, so for
stft/istft the published algorithm is restated (oracle/transform.py) and
cross-checked against torch.stft/istft; that one leg is "parity unpinned"
against librosa's own binaries.
"""
