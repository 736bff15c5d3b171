"""Stage A oracle: dual (slow/fast) autoregressive decode.  TEST INFRASTRUCTURE ONLY.

Plain torch-fp32 CPU restatement of `ARVCWrapper` (modules/arvc_wrapper.py:82-126),
`DualARWrapper` (modules/dual_ar_stream.py:698-837), `decode_one_token_ar` (:1168-1219)
and the sampler (:1081-1132) for configs/hydra_arcs/vc/firefly_arvc_bsq_8192_delay0_8.yaml.
KV caches are fp32 (the shipped fp16 cache cannot run on CPU, SURVEY.md finding 1).

Sampling noise is an explicit argument: `noise_fn(step, slot, V)` returns the Exp(1)
vector `q` that `multinomial_sample_one_no_sync` (:1092-1096) would have drawn for the
`slot`-th sampler call (0 = 8192-way token head, 1..8 = 1000-way codebooks) of the
`step`-th `decode_one_token_ar` call since `reset_steps()`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

DIM = 768
N_HEAD = 12
HEAD_DIM = 64
N_LAYER = 12
N_FAST = 4
N_CB = 8
CB_SIZE = 1000
VOCAB = 8192
MAX_SEQ = 2048
EPS = 1e-5


def rope_table(seq_len, n_elem=HEAD_DIM, base=10000.0):
    """precompute_freqs_cis, dual_ar_stream.py:993-1001 (bf16-rounded cos/sin)."""
    freqs = 1.0 / (base ** (torch.arange(0, n_elem, 2)[: n_elem // 2].float() / n_elem))
    freqs = torch.outer(torch.arange(seq_len), freqs)
    cis = torch.polar(torch.ones_like(freqs), freqs)
    return torch.stack([cis.real, cis.imag], dim=-1).to(torch.bfloat16)


def apply_rope(x, fc):
    """apply_rotary_emb, dual_ar_stream.py:1004-1016.  x [B,S,H,D]; fc [B,S,D/2,2]."""
    xs = x.float().reshape(*x.shape[:-1], -1, 2)
    fc = fc.view(x.size(0), xs.size(1), 1, xs.size(3), 2)
    out = torch.stack([xs[..., 0] * fc[..., 0] - xs[..., 1] * fc[..., 1],
                       xs[..., 1] * fc[..., 0] + xs[..., 0] * fc[..., 1]], -1)
    return out.flatten(3).type_as(x)


def rms_norm(x, w):
    """RMSNorm.forward, dual_ar_stream.py:979-990."""
    xf = x.float()
    return (xf * torch.rsqrt(torch.mean(xf * xf, dim=-1, keepdim=True) + EPS)).type_as(x) * w


def logits_to_probs(logits, temperature=0.7, top_p=0.7):
    """dual_ar_stream.py:1099-1132 without repetition penalty (previous_tokens is None on
    every call site of the hot path)."""
    sorted_logits, sorted_indices = torch.sort(logits, descending=True)
    cum_probs = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
    remove = cum_probs > top_p
    remove[0] = False
    remove = remove.scatter(dim=0, index=sorted_indices, src=remove)
    logits = logits.masked_fill(remove, -float("Inf"))
    logits = logits / max(temperature, 1e-5)
    return F.softmax(logits, dim=-1)


def sample(logits, q, **sampling_kwargs):
    """sample + multinomial_sample_one_no_sync, dual_ar_stream.py:1081-1096, with the
    Exp(1) draw `q` supplied by the caller."""
    probs = logits_to_probs(logits[0, -1], **sampling_kwargs)
    return torch.argmax(probs / q[: probs.shape[0]], dim=-1, keepdim=True).to(torch.int)


class DualAR:
    """Holds weights (flat reference state-dict), KV caches and the streaming state that
    the reference keeps inside DualARWrapper (dual_ar_stream.py:775-796,808-815,834-836)."""

    def __init__(self, sd, noise_fn=None, max_seq_len=MAX_SEQ):
        self.sd = sd
        self.max_seq_len = max_seq_len
        self.freqs = rope_table(MAX_SEQ)
        self.fast_freqs = rope_table(N_CB)
        self.causal = torch.tril(torch.ones(MAX_SEQ, MAX_SEQ, dtype=torch.bool))
        self.k = torch.zeros(N_LAYER, 1, N_HEAD, max_seq_len, HEAD_DIM)
        self.v = torch.zeros_like(self.k)
        self.fk = torch.zeros(N_FAST, 1, N_HEAD, N_CB, HEAD_DIM)
        self.fv = torch.zeros_like(self.fk)
        self.delay = 0
        self.noise_fn = noise_fn
        self.step = 0
        self.last_logits = None      # teacher-forced diagnostics: [9] list of logits of the last call
        self.last_hidden = None

    # ------------------------------------------------------------------ model pieces
    def embed_codes(self, codes):
        """BaseTransformer.embed, dual_ar_stream.py:245-255.  codes [B,8,T] -> [B,T,768]."""
        tab = self.sd["decoder.model.codebook_embeddings.weight"]
        embs = [F.embedding(codes[:, i].long() + i * CB_SIZE, tab) for i in range(N_CB)]
        return torch.stack(embs, dim=3).sum(dim=3)

    def _block(self, x, p, fc, mask, kc, vc, pos):
        """TransformerBlock.forward :854-861, Attention.forward :895-936 (no cross-attn),
        KVCache.update :141-150, FeedForward :975-976."""
        sd = self.sd
        B, S, _ = x.shape
        h = rms_norm(x, sd[p + ".attention_norm.weight"])
        q, k, v = F.linear(h, sd[p + ".attention.wqkv.weight"]).split([DIM, DIM, DIM], dim=-1)
        q = apply_rope(q.view(B, S, N_HEAD, HEAD_DIM), fc).transpose(1, 2)
        k = apply_rope(k.view(B, S, N_HEAD, HEAD_DIM), fc).transpose(1, 2)
        v = v.view(B, S, N_HEAD, HEAD_DIM).transpose(1, 2)
        kc[:, :, pos] = k
        vc[:, :, pos] = v
        y = F.scaled_dot_product_attention(q, kc, vc, attn_mask=mask)
        y = y.transpose(1, 2).contiguous().view(B, S, DIM)
        h = x + F.linear(y, sd[p + ".attention.wo.weight"])
        n = rms_norm(h, sd[p + ".ffn_norm.weight"])
        f = F.linear(F.silu(F.linear(n, sd[p + ".feed_forward.w1.weight"])) * F.linear(n, sd[p + ".feed_forward.w3.weight"]),
                     sd[p + ".feed_forward.w2.weight"])
        return h + f

    def forward_generate(self, x, input_pos, kv_pos):
        """BaseTransformer.forward_generate, dual_ar_stream.py:312-356: returns the token
        logits of the last position and its PRE-norm hidden state."""
        mask = self.causal[None, None, kv_pos, : self.max_seq_len]
        fc = self.freqs[input_pos]
        for i in range(N_LAYER):
            x = self._block(x, f"decoder.model.layers.{i}", fc, mask, self.k[i], self.v[i], kv_pos)
        x = x[:, -1:]
        logits = F.linear(rms_norm(x, self.sd["decoder.model.norm.weight"]), self.sd["decoder.model.output.weight"])
        return logits, x

    def forward_generate_fast(self, x, cb):
        """DualARTransformer.forward_generate_fast, dual_ar_stream.py:540-558."""
        pos = torch.tensor([cb])
        x = x.view(1, 1, -1)
        mask = self.causal[None, None, pos, :N_CB]
        fc = self.fast_freqs[pos]
        for i in range(N_FAST):
            x = self._block(x, f"decoder.model.fast_layers.{i}", fc[None], mask, self.fk[i], self.fv[i], pos)
        return F.linear(rms_norm(x, self.sd["decoder.model.fast_norm.weight"]), self.sd["decoder.model.fast_output.weight"])

    def decode_one_token_ar(self, x, input_pos, kv_pos, **sampling_kwargs):
        """dual_ar_stream.py:1168-1219.  Returns int32 [9,1] (token-head sample + 8 codes)."""
        step = self.step
        self.step += 1
        logits, hidden = self.forward_generate(x, input_pos, kv_pos)
        all_logits = [logits[0, -1].clone()]
        codebooks = [sample(logits, self.noise_fn(step, 0, VOCAB), **sampling_kwargs)]
        self.fk.zero_()
        self.fv.zero_()
        self.last_hidden = hidden[0, 0].clone()
        for cb in range(N_CB):
            logits = self.forward_generate_fast(hidden, cb)
            all_logits.append(logits[0, -1].clone())
            a = sample(logits, self.noise_fn(step, cb + 1, CB_SIZE), **sampling_kwargs)
            hidden = F.embedding(a, self.sd["decoder.model.fast_embeddings.weight"])
            codebooks.append(a.clone())
        self.last_logits = all_logits
        return torch.stack(codebooks, dim=0)

    # ------------------------------------------------------------------ ARVCWrapper surface
    def set_delay(self, delay):
        self.delay = int(delay)

    def _cond(self, content_codes):
        return F.embedding(content_codes.long(), self.sd["embedding.weight"])

    def _spk(self, style_vectors, timbre_latents):
        """arvc_wrapper.py:108-109: [context_in(timbre) (32 tokens) | style_in(style) (1 token)]."""
        sd = self.sd
        return torch.cat([F.linear(timbre_latents, sd["context_in.weight"], sd["context_in.bias"]),
                          F.linear(style_vectors, sd["style_in.weight"], sd["style_in.bias"]).unsqueeze(1)], dim=1)

    def prefill_prompt(self, ref_content_codes, ref_audio_codes, style_vectors, timbre_latents):
        """arvc_wrapper.py:100-112 -> dual_ar_stream.py:764-796."""
        d = self.delay
        ref_cond = self._cond(ref_content_codes)
        spk = self._spk(style_vectors, timbre_latents)
        B, T, D = ref_cond.shape
        ref_emb = self.embed_codes(ref_audio_codes)
        self.cached_ref_emb = ref_emb[:, -d:].clone() if d != 0 else ref_emb
        if d != 0:
            w4s = self.sd["decoder.wait4start_embedding.weight"][:d]
            ref_emb = torch.cat([w4s.unsqueeze(0), ref_emb[:, :-d]], dim=1)
        else:
            self.cached_new_audio_emb = ref_emb[:, -1:].clone()
        emb_seq = torch.stack([ref_cond, ref_emb], dim=1).transpose(1, 2).reshape(B, -1, D)
        emb_seq = torch.cat([spk, emb_seq], dim=1)
        if d == 0:
            emb_seq = emb_seq[:, :-1]
        input_pos = torch.arange(emb_seq.size(1))[None]
        kv_pos = torch.arange(emb_seq.size(1))
        self.decode_one_token_ar(emb_seq, input_pos, kv_pos)
        self.cached_input_pos, self.cached_kv_pos = input_pos, kv_pos

    def prefill_src_condition4delay(self, src_content_codes):
        """arvc_wrapper.py:114-119 -> dual_ar_stream.py:798-815."""
        src_cond = self._cond(src_content_codes)
        assert src_cond.size(1) == self.delay
        B, T, D = src_cond.shape
        emb_seq = torch.stack([src_cond, self.cached_ref_emb], dim=1).transpose(1, 2).reshape(B, -1, D)
        self.cached_new_audio_emb = emb_seq[:, -1:].clone()
        emb_seq = emb_seq[:, :-1]
        input_pos = torch.arange(emb_seq.size(1))[None] + self.cached_input_pos[:, -1:] + 1
        kv_pos = torch.arange(emb_seq.size(1)) + self.cached_kv_pos[-1:] + 1
        self.decode_one_token_ar(emb_seq, input_pos, kv_pos)
        self.cached_input_pos, self.cached_kv_pos = input_pos, kv_pos

    def decode_one(self, src_content_codes):
        """arvc_wrapper.py:121-126 -> dual_ar_stream.py:817-837.  Returns (int32 [8,1], last pos)."""
        src_cond = self._cond(src_content_codes)
        emb_seq = torch.cat([self.cached_new_audio_emb, src_cond], dim=1)
        input_pos = torch.arange(2)[None] + self.cached_input_pos[:, -1:] + 1
        kv_pos = torch.arange(2) + self.cached_kv_pos[-1:] + 1
        nxt = self.decode_one_token_ar(emb_seq, input_pos, kv_pos)
        pred_x = nxt[1:]
        self.cached_new_audio_emb = self.embed_codes(pred_x.unsqueeze(0))
        self.cached_input_pos, self.cached_kv_pos = input_pos, kv_pos
        return pred_x, kv_pos[-1]

    def generate(self, ref_content_codes, ref_audio_codes, src_content_codes, style_vectors, timbre_latents,
                 **sampling_kwargs):
        """arvc_wrapper.py:82-98 -> dual_ar_stream.py:698-762 (offline).  The first frame is
        sampled with the DEFAULT sampling arguments (:723 passes no kwargs).  -> [1,8,Ts]."""
        d = self.delay
        src_cond = self._cond(src_content_codes)
        ref_cond = self._cond(ref_content_codes)
        spk = self._spk(style_vectors, timbre_latents)
        B, T, D = ref_cond.shape
        w4e = self.sd["decoder.wait4end_embedding.weight"][:d]
        w4s = self.sd["decoder.wait4start_embedding.weight"][:d]
        ref_emb = torch.cat([w4s.unsqueeze(0), self.embed_codes(ref_audio_codes)], dim=1)
        prefill_cond = torch.cat([ref_cond, src_cond[:, :d]], dim=1)
        emb_seq = torch.stack([prefill_cond, ref_emb], dim=1).transpose(1, 2).reshape(B, -1, D)
        emb_seq = torch.cat([spk, emb_seq], dim=1)
        remaining = torch.cat([src_cond[:, d:], w4e.unsqueeze(0)], dim=1)
        emb_seq = torch.cat([emb_seq, remaining[:, :1]], dim=1)
        input_pos = torch.arange(emb_seq.size(1))[None]
        kv_pos = torch.arange(emb_seq.size(1))
        pred = self.decode_one_token_ar(emb_seq, input_pos, kv_pos)
        pred_codes = [pred[1:]]
        for i in range(remaining.size(1) - 1):
            new_audio = self.embed_codes(pred_codes[-1].unsqueeze(0))
            emb_seq = torch.cat([new_audio, remaining[:, i + 1 : i + 2]], dim=1)
            input_pos = input_pos[:, -2:] + 2
            kv_pos = kv_pos[-2:] + 2
            nxt = self.decode_one_token_ar(emb_seq, input_pos, kv_pos, **sampling_kwargs)
            pred_codes.append(nxt[1:].clone())
        return torch.stack(pred_codes, dim=-1).transpose(0, 1)
