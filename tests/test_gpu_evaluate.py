"""GPU parity of the sharded batch evaluation (BASELINE config 4, scripts/evaluate_AV_net.py): a batched call with the
per-utterance MCB norm must return what the reference returns when it calls the model once per utterance."""
import numpy as np
import pytest
import torch

from avvad import engine as E
from avvad import synth
from avvad.evaluate import evaluate_shard, evaluate_sharded, plan_calls
from avvad.pipeline import AVVADPipeline
from oracle.reference_port import RefDeepVADAV, cpu_av_step
from util import check_logits, eval_single_inputs, golden

pytestmark = pytest.mark.gpu


def _utterances(n, seed=0):
    ns = [16000 + 2311 * ((5 * i + 3) % n) for i in range(n)]     # ragged, not sorted
    nf = [30 + 4 * ((5 * i + 3) % n) for i in range(n)]
    return [(synth.synth_wave(a, seed + i), synth.synth_video_u8(f, seed + i)) for i, (a, f) in enumerate(zip(ns, nf))]


def _logit(p):
    p = np.clip(np.asarray(p, dtype=np.float64), 1e-12, 1.0 - 1e-12)
    return np.log(p) - np.log1p(-p)


def test_grouped_mcb_equals_one_call_per_utterance():
    """avvad_mcb_forward_grouped on (B, t_max) rows == avvad_mcb_forward on each utterance's valid rows alone; rows behind
    an utterance's length are zeros.  Includes an utterance of length 0 and one that fills t_max."""
    g = torch.Generator().manual_seed(3)
    B, t_max = 5, 23
    lens = [23, 1, 0, 17, 8]
    audio = torch.randn(B, t_max, 513, generator=g).cuda()
    video = torch.randn(B, t_max, 512, generator=g).abs().cuda()
    sd = synth.calibrate_mcb_bn_(synth.seeded_state_dict(synth.model_spec("av", use_mcb=True), 9), 16)
    mcb = E.Mcb()
    mcb.load(sd, "cuda")
    out_bf = torch.full((B, t_max, 1024), 7.0, dtype=torch.bfloat16, device="cuda")
    out32 = torch.full((B, t_max, 1024), 7.0, dtype=torch.float32, device="cuda")
    mcb.forward_grouped(audio, video, lens, t_max, out_bf16=out_bf.view(-1, 1024), out_f32=out32.view(-1, 1024))
    for b, n in enumerate(lens):
        assert torch.count_nonzero(out32[b, n:]) == 0 and torch.count_nonzero(out_bf[b, n:].float()) == 0
        if n == 0:
            continue
        one = torch.empty(n, 1024, dtype=torch.float32, device="cuda")
        mcb.forward(audio[b, :n], video[b, :n], out_f32=one)
        # same row kernel; the two norm kernels add the same fp32 row sums in a different fp64 order
        assert torch.allclose(out32[b, :n], one, rtol=2e-6, atol=1e-7), (b, (out32[b, :n] - one).abs().max())
        assert torch.equal(out_bf[b, :n], out32[b, :n].to(torch.bfloat16))


def test_batched_evaluation_matches_reference_single_utterance_calls():
    """evaluate_shard (one batched, length-sorted device call) vs the CPU port of the reference called with ONE utterance
    per forward, as scripts/evaluate_AV_net.py:186-236 does.  Strong weight family: logits span several units."""
    utts = _utterances(6, seed=11)
    mean, std = synth.synth_audio_stats(0)
    sd = synth.calibrate_mcb_bn_(synth.seeded_state_dict(synth.model_spec("av", use_mcb=True), 51, "strong"), 90)
    ref_model = RefDeepVADAV(2, 1024, 1, use_mcb=True).load_reference_state_dict(sd).eval()
    ref_logit = []
    for w, v in utts:
        post, _, lens = cpu_av_step(ref_model, [w], [v], mean, std, synth.VIDEO_MEAN, synth.VIDEO_STD)
        ref_logit.append(_logit(post[0, : lens[0]]))
    # place the head bias so that the bulk of the logits sits 2.5 sigma off the threshold (synth.decision_bias; logits are
    # affine in the bias, so the reference needs no second pass)
    allref = np.concatenate(ref_logit)
    bias = sd["vad_merged.bias"].numpy().astype(np.float64)
    nb = synth.decision_bias(allref[None, :, None], [allref.size], bias)
    ref_logit = [r - bias[0] + float(nb[0]) for r in ref_logit]
    sd["vad_merged.bias"] = nb
    pipe = AVVADPipeline(sd, mean, std, synth.VIDEO_MEAN, synth.VIDEO_STD, use_mcb=True)
    res = evaluate_shard(pipe, utts, batch_size=4)       # two calls: 4 + 2 utterances, sorted by length
    assert len(res) == len(utts)
    got, want = [], []
    for (soft, hard), r in zip(res, ref_logit):
        assert soft.shape == (len(r),) and hard.shape == (len(r),)
        assert torch.equal(hard, (soft > 0.5).to(hard.dtype))
        got.append(_logit(soft))
        want.append(r)
    got, want = np.concatenate(got), np.concatenate(want)
    check_logits(got[None, :, None], want[None, :, None], None, "config 4: batched vs one call per utterance")


def test_posteriors_do_not_depend_on_the_call_grouping():
    """Size-independent property of the per-utterance mode: any grouping / order / padding of the same utterances gives
    bit-identical posteriors (per-frame trunk, per-row MCB, per-utterance norm, per-row recurrence), which is what lets
    the shard be sorted by length.  Also: rank blocks of evaluate_sharded tile the list."""
    utts = _utterances(7, seed=23)
    mean, std = synth.synth_audio_stats(0)
    sd = synth.calibrate_mcb_bn_(synth.seeded_state_dict(synth.model_spec("av", use_mcb=True), 51, "strong"), 90)
    pipe = AVVADPipeline(sd, mean, std, synth.VIDEO_MEAN, synth.VIDEO_STD, use_mcb=True)
    a = evaluate_shard(pipe, utts, batch_size=7, sort_by_length=True)     # one call
    b = evaluate_shard(pipe, utts, batch_size=3, sort_by_length=False)    # 3 + 3 + 1 in list order
    c = evaluate_shard(pipe, utts, batch_size=1)                          # one utterance per call, no padding at all
    for (sa, ha), (sb, hb), (sc, hc) in zip(a, b, c):
        assert torch.equal(sa, sb) and torch.equal(ha, hb)
        assert torch.equal(sa, sc) and torch.equal(ha, hc)
    got = {}
    for rank in range(3):
        lo, hi, res = evaluate_sharded(pipe, utts, 3, rank, batch_size=2)
        for i, r in zip(range(lo, hi), res):
            got[i] = r
    assert sorted(got) == list(range(len(utts)))
    for i, (s, h) in got.items():
        assert torch.equal(s, a[i][0]) and torch.equal(h, a[i][1])
    # the sorted plan starts with the longest utterance
    n = [int(np.asarray(w).shape[-1]) for w, _ in utts]
    f = [int(np.asarray(v).shape[0]) for _, v in utts]
    T = AVVADPipeline.frame_counts(n, f)
    assert plan_calls(T, 7)[0][0] == int(np.argmax(T))


def test_batched_module_call_matches_reference_called_once_per_utterance():
    """Pinned against the reference itself: tests/golden/ref_eval_single.npz holds the logits of the UNMODIFIED
    DeepVAD_AV(use_mcb=True) called once per utterance (scripts/evaluate_AV_net.py:186-236).  ONE batched forward of the
    drop-in module with norm_per_utterance = True must reproduce them; with the flag off it reproduces the reference's
    batched call instead (whole-call norm over the padded tensor), which differs by up to 1.1 in the logits."""
    from packages.models.AV_Net import DeepVAD_AV
    g = golden("ref_eval_single.npz")
    a, v, lens = eval_single_inputs()
    sd = synth.calibrate_mcb_bn_(synth.seeded_state_dict(synth.model_spec("av", use_mcb=True), 43, "strong"), 20)
    sd["vad_merged.bias"] = torch.tensor(g["bias"])
    m = DeepVAD_AV(2, 1024, 1, use_mcb=True, eps=1e-8)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.norm_per_utterance = True
    out = m(torch.tensor(a).cuda(), torch.tensor(v).cuda(), lens).cpu().numpy()
    check_logits(out, g["logits"], lens, "module, per-utterance norm vs reference single calls")
    m.norm_per_utterance = False
    out = m(torch.tensor(a).cuda(), torch.tensor(v).cuda(), lens).cpu().numpy()
    # (the head bias was placed for the single-call logits, so only the two error gates here, not the decision gate)
    mask = np.zeros(out.shape[:2], dtype=bool)
    for b, n in enumerate(lens):
        mask[b, :n] = True
    ref = g["batched_call_logits"][mask].astype(np.float64)
    got = out[mask].astype(np.float64)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 2e-2
    assert np.abs(1 / (1 + np.exp(-got)) - 1 / (1 + np.exp(-ref))).max() <= 1e-2
