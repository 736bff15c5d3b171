"""Import shim (test infrastructure only) for the single einx pattern the reference's
vendored FSQ uses (modules/bicodec_speaker_encoder/fsq/residual_fsq.py:136):
get_at("q [c] d, b n q -> q b n d", codebooks, indices)."""
import torch


def get_at(pattern, codebooks, indices):
    assert pattern.replace(" ", "") == "q[c]d,bnq->qbnd", pattern
    q = codebooks.shape[0]
    out = [codebooks[i][indices[..., i]] for i in range(q)]
    return torch.stack(out, dim=0)
