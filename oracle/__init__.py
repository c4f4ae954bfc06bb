"""CPU oracle for the AV-VAD hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy / fp32 PyTorch-on-CPU, the algorithms that
sp-uhh/audio-visual-vad runs on the path named by BASELINE.json (front end -> upsample ->
ResNet-18 -> concat/MCB -> LSTM -> head, plus loss/metrics).  Every function cites the reference
file:line it follows.  It exists so that the CUDA path can be checked for parity; it is never the
product:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import it;
  * nothing under ``audio-visual-vad_b200/`` imports it, and the product path raises when the
    CUDA library is missing instead of falling back to this code.

Pinning status (details in DESIGN.md "Oracle"):
  * front end / pad rule / frame counts, IBM + VAD labels, upsampling index map, DCT->ROI decode:
    pinned against the reference's own golden files (tests/golden/*.npz, made by
    tools/make_golden.py from /root/reference/data/subset).
  * Audio / Video / AV(concat) forward, count sketch, losses, collate, WaveNet encoder: pinned
    against outputs of the reference's own modules imported in the build container
    (tests/golden/ref_*.npz, same script).
  * AV with use_mcb=True: the reference cannot execute on torch>=1.8 (torch.rfft removed); pinned
    piecewise (reference CountSketchFn_forward + circular-convolution identity), the rest of that
    branch is a restatement only.
"""
