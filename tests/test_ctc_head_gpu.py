"""The recognizer head + CTC greedy decode on its own (SURVEY §8 a-R2, reference src/ocr_rec.cpp:97-128), through
b200ocr_kernel_ctc_head: given IDENTICAL upstream tensors (features and weights that are exactly representable in
fp16, which is what the device holds) the decoded indices must be BIT-EXACT against a torch fp32 arg-max +
oracle.ocr_ops.ctc_collapse, on both the tcgen05 kernel and the CUDA-core kernel.

Two kinds of input:
 * exact arithmetic: features k/8 and weights k/16 with small integers k and 120 channels, so that every product and
   every partial sum is exactly representable in fp32 -- the logits do not depend on the summation order, ties are
   real ties, and "the first maximum wins" (Utility::argmax = std::max_element) is checked bit for bit;
 * Gaussian features / weights rounded to fp16: the logits then differ from torch's by fp32 summation order only
   (~1e-6 relative), so indices must be equal wherever the fp32 top-2 logit gap exceeds 1e-4; the soft-max probability
   of the arg-max (the reference's max_value) must agree within 1e-3 absolute (north star: 1e-2).
"""
import numpy as np
import pytest

from oracle import ocr_ops as ops

pytestmark = pytest.mark.gpu

NCLS = 6625
CIN = 120


def _h(a):
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def _torch_head(feat, w, bias):
    import torch
    logits = torch.tensor(feat, dtype=torch.float32) @ torch.tensor(w, dtype=torch.float32) + torch.tensor(bias)
    probs = torch.softmax(logits, -1)
    return logits.numpy(), probs.numpy()


def _collapse_ref(idx_row, mx_row):
    labels = list(range(NCLS))  # label == id
    out, score, count, last = [], np.float32(0), 0, 0
    for t in range(len(idx_row)):
        a = int(idx_row[t])
        if a > 0 and not (t > 0 and a == last):
            out.append(a)
            score = np.float32(score + np.float32(mx_row[t]))
            count += 1
        last = a
    return out, (np.float32(score / np.float32(count)) if count else np.float32(0))


@pytest.mark.parametrize("simt", [False, True], ids=["tcgen05", "cuda-core"])
@pytest.mark.parametrize("n,t", [(1, 24), (5, 40), (16, 97), (3, 129)])
def test_ctc_head_indices_bit_exact_on_exact_arithmetic(simt, n, t):
    import b200ocr
    rng = np.random.default_rng(n * 1000 + t)
    feat = rng.integers(-6, 7, (n, t, CIN)).astype(np.float32) / 8.0
    w = rng.integers(-8, 9, (CIN, NCLS)).astype(np.float32) / 16.0
    # force real ties and repeats: duplicate weight columns (the lower class id must win), blank-heavy and repeated steps
    w[:, 3000:6000] = w[:, 0:3000]  # class k + 3000 duplicates class k (the blank included): a winner below 6000 ties
    w[:, 6624] = w[:, 300]
    feat[:, 1::5] = feat[:, 0::5][:, : feat[:, 1::5].shape[1]]  # step t+1 == step t -> same arg-max -> collapsed
    bias = np.zeros(NCLS, np.float32)
    logits, probs = _torch_head(feat, w, bias)
    ref_idx = logits.argmax(-1)  # numpy arg-max: first maximum wins, like std::max_element
    assert (np.sort(logits, -1)[..., -1] == np.sort(logits, -1)[..., -2]).any(), "the case is meant to contain exact ties"
    idx, prob, collapsed, scores = b200ocr.kernel_ctc_head(feat, w, bias, force_simt=simt)
    assert np.array_equal(idx, ref_idx.astype(np.int32)), int((idx != ref_idx).sum())
    ref_mx = probs.max(-1)
    assert float(np.abs(prob - ref_mx).max()) <= 1e-3
    for i in range(n):
        want, want_score = _collapse_ref(ref_idx[i], prob[i])  # collapse rule on identical (idx, prob): bit-exact
        assert list(collapsed[i]) == want
        assert np.float32(scores[i]) == want_score, (scores[i], want_score)
        # and the oracle's own decode of the torch probabilities gives the same ids
        labels = [str(k) + "," for k in range(NCLS)]
        r = ops.ctc_greedy_decode(probs[i], labels)
        assert (r[0] if r else "") == "".join(labels[k] for k in want)


@pytest.mark.parametrize("simt", [False, True], ids=["tcgen05", "cuda-core"])
def test_ctc_head_on_gaussian_features(simt):
    import b200ocr
    rng = np.random.default_rng(7)
    n, t = 64, 97
    feat = _h(rng.standard_normal((n, t, CIN)))
    w = _h(rng.standard_normal((CIN, NCLS)) * (4.0 / np.sqrt(CIN)))
    bias = (rng.standard_normal(NCLS) * 0.5).astype(np.float32)
    logits, probs = _torch_head(feat, w, bias)
    srt = np.sort(logits, -1)
    decided = (srt[..., -1] - srt[..., -2]) > 1e-4
    idx, prob, collapsed, scores = b200ocr.kernel_ctc_head(feat, w, bias, force_simt=simt)
    ref_idx = logits.argmax(-1)
    assert decided.mean() > 0.999
    assert np.array_equal(idx[decided], ref_idx[decided].astype(np.int32))
    assert float(np.abs(prob - probs.max(-1))[decided].max()) <= 1e-3
    for i in range(n):
        want, want_score = _collapse_ref(idx[i], prob[i])
        assert list(collapsed[i]) == want and np.float32(scores[i]) == want_score
