"""Prompt-context caching for LeftRefill's learned-prompt text encoders.

The reference re-runs its frozen OpenCLIP ViT-H/14 text tower on every batch (ldm/modules/encoders/Refill_modules.py:
PromptCLIPEmbedder.forward :160-178 -> encode_with_transformer :180-191, called through
LatentDiffusion.get_learned_conditioning, ddpm.py:677-690, from get_input and from
RefInpaintLDM.get_unconditional_conditioning, inpainting_ldm/ref_inpainting_ldm.py:30-35,42) although the prompt is a
constant string per config (dataloaders/test_dataset.py:39-49) and the unconditional prompt is always "". The context
`[B, 77, 1024]` is therefore a constant of the process: a 24-layer text forward per batch buys nothing.

`PromptContextCache` wraps any such embedder (a callable `text -> [B, L, C]` tensor, or `[B, n_layer, L, C]` for the
deep-prompt variant that receives a list of per-layer prompt lists) and memoises the result per UNIQUE prompt string:
the embedder runs once per new string (batched over the new strings only), every later batch is an index_select over
the cached rows. The cache is dropped when a parameter of the embedder changes (storage / version / dtype / device
signature, plus an explicit `invalidate()` for `.data` writes), so training the learned special-token embeddings
(`special_embeddings`, the only trainable part, Refill_modules.py:134-140) never sees stale rows.

`install_context_cache(ldm)` swaps `ldm.cond_stage_model` for the wrapper, leaving `get_learned_conditioning` and the
drivers unchanged. This is host-side plumbing around a frozen encoder: no arithmetic of the encoder is re-implemented.
"""
import torch
import torch.nn as nn


class PromptContextCache(nn.Module):
    def __init__(self, embedder, max_entries=4096):
        super().__init__()
        self.embedder = embedder
        self.max_entries = int(max_entries)
        self._rows = {}      # prompt key -> cached tensor for ONE sample ([L, C] or [n_layer, L, C])
        self._sig = None
        self.hits = 0
        self.misses = 0
        self.encoder_calls = 0

    # the reference reaches the embedder through either name (ddpm.py:679-686)
    def encode(self, text):
        return self(text)

    def invalidate(self):
        self._rows.clear()
        self._sig = None

    def _signature(self):
        if not isinstance(self.embedder, nn.Module):
            return None
        return tuple((p.data_ptr(), p._version, p.dtype, str(p.device)) for p in self.embedder.parameters())

    @staticmethod
    def _is_deep(text):
        return len(text) > 0 and isinstance(text[0], (list, tuple))

    def _call(self, text):
        self.encoder_calls += 1
        enc = getattr(self.embedder, "encode", None)
        # PromptCLIPEmbedder defines both; FrozenOpenCLIPEmbedder-style encoders only `encode` -> forward
        return self.embedder(text) if callable(self.embedder) else enc(text)

    def forward(self, text):
        if isinstance(text, str):
            text = [text]
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._call(text)  # training the prompt embeddings: never serve cached (detached) rows
        with torch.no_grad():
            return self._cached(text)

    def _cached(self, text):
        sig = self._signature()
        if sig != self._sig:
            self._rows.clear()
            self._sig = sig
        deep = self._is_deep(text)
        # deep prompt: text is [n_layer][B] (ref_inpainting_ldm.py:32); a sample's key is its column of prompts
        keys = [tuple(layer[i] for layer in text) for i in range(len(text[0]))] if deep else list(text)
        new = [k for k in dict.fromkeys(keys) if k not in self._rows]
        self.misses += len(new)
        self.hits += len(keys) - len(new)
        if new:
            batch = [[k[j] for k in new] for j in range(len(text))] if deep else new
            z = self._call(batch)
            assert z.shape[0] == len(new), "embedder returned an unexpected batch size"
            if len(self._rows) + len(new) > self.max_entries:
                self._rows.clear()
            for i, k in enumerate(new):
                self._rows[k] = z[i].detach().clone()
        uniq = list(dict.fromkeys(keys))
        table = torch.stack([self._rows[k] for k in uniq])
        if len(uniq) == len(keys):
            return table
        pos = {k: i for i, k in enumerate(uniq)}
        idx = torch.tensor([pos[k] for k in keys], device=table.device)
        return table.index_select(0, idx)

    def __getattr__(self, name):  # tokenizer, special_tokens, device ... stay reachable through the wrapper
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("embedder"), name)


def install_context_cache(ldm, max_entries=4096):
    """Wraps `ldm.cond_stage_model` (LatentDiffusion.instantiate_cond_stage, ddpm.py:626-653) in a PromptContextCache and
    returns the wrapper. Idempotent."""
    m = ldm.cond_stage_model
    if isinstance(m, PromptContextCache):
        return m
    w = PromptContextCache(m, max_entries=max_entries)
    ldm.cond_stage_model = w
    return w
