"""Greedy decoding + OCR metrics for the eval path.

Mirrors ``pixparse.utils.ocr_utils`` (/root/reference/src/pixparse/utils/ocr_utils.py:15-222): encoder once per
batch, then the UNCACHED greedy loop the reference runs -- the whole prefix is re-fed every step
(``prepare_inputs_for_inference`` with ``past_key_values=None``), finished rows keep generating, the loop stops when
every row has emitted EOS or at ``max_recursion_length``. CER / WER are computed with a local Levenshtein
implementation (jiwer is not a dependency of the hot path).
"""
import re
from typing import List

import torch


def _edit_distance(a, b):
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def _clean(text):
    return re.sub(r"\s+", " ", text.replace("<pad>", "")).strip()


def cer(reference: List[str], hypothesis: List[str]) -> float:
    errs = sum(_edit_distance(list(_clean(r)), list(_clean(h))) for r, h in zip(reference, hypothesis))
    total = sum(len(_clean(r)) for r in reference)
    return errs / max(total, 1)


def wer(reference: List[str], hypothesis: List[str]) -> float:
    errs = sum(_edit_distance(_clean(r).split(), _clean(h).split()) for r, h in zip(reference, hypothesis))
    total = sum(len(_clean(r).split()) for r in reference)
    return errs / max(total, 1)


def get_next_token(next_token_logits, use_sample: bool = True, temperature: float = 5):
    if use_sample:
        probs = torch.nn.functional.softmax(next_token_logits.float() / temperature, dim=-1)
        next_token_id = torch.multinomial(probs, num_samples=1).reshape(-1).unsqueeze(-1)
    else:
        next_token_id = next_token_logits.argmax(1).unsqueeze(-1)
        probs = torch.ones_like(next_token_logits)
    return next_token_id, probs


def get_generated_tokens(model, tokenizer, encoder_outputs, device_env, max_recursion_length, prompt_token: str,
                         use_cache: bool = False, stop_on_eos: bool = True, graph_decode: bool = True):
    """use_cache=False is the reference loop (whole prefix re-fed each step). use_cache=True feeds only the last token
    and carries ``past_key_values`` (the branch prepare_inputs_for_inference already has, text_decoder_hf.py:69-70):
    same token ids, O(steps) instead of O(steps^2) decoder work; with graph_decode (default, <= 16 pages) the cached loop
    runs as one replayed CUDA graph per token, otherwise through TextDecoderHf.forward(past_key_values=...).
    stop_on_eos=False (benchmarks with random weights) always runs max_recursion_length steps and never reads the
    "all rows finished" flag back to the host."""
    task_prompt_id = tokenizer.trunk.encode(prompt_token, add_special_tokens=False)[0]
    device = device_env.device
    if use_cache and graph_decode and encoder_outputs.shape[0] <= 16:
        # up to 16 pages: the whole loop below as a replayed CUDA graph of single-token kernels (pixparse_b200/decode.py)
        from .engine import engine_for
        return engine_for(model.text_decoder).greedy_decode(
            encoder_outputs, task_prompt_id, max_recursion_length, tokenizer.trunk.eos_token_id,
            tokenizer.trunk.pad_token_id, stop_on_eos=stop_on_eos)
    input_ids = torch.full((encoder_outputs.shape[0], 1), task_prompt_id, dtype=torch.long, device=device)
    finished = torch.zeros(input_ids.shape[0], dtype=torch.bool, device=device)
    eos_token_id = tokenizer.trunk.eos_token_id
    past = None
    for _ in range(max_recursion_length):
        inputs = model.text_decoder.prepare_inputs_for_inference(
            input_ids=input_ids, encoder_outputs=encoder_outputs, pad_token_id=tokenizer.trunk.pad_token_id,
            past_key_values=past, use_cache=True if use_cache else None)
        outputs = model.text_decoder.forward(**inputs)
        if use_cache:
            past = outputs.past_key_values
        next_token_logits = outputs.logits[:, -1, :]
        next_token_id, _ = get_next_token(next_token_logits, use_sample=False)
        if stop_on_eos:
            finished |= next_token_id.squeeze(-1) == eos_token_id
            if finished.all():
                break
        input_ids = torch.cat([input_ids, next_token_id], dim=-1)
    return input_ids


def generate_ocr(model, tokenizer, encoder_outputs, device_env, max_recursion_length, prompt_token: str) -> List[str]:
    with torch.inference_mode():
        tokens = get_generated_tokens(model, tokenizer, encoder_outputs, device_env, max_recursion_length, prompt_token)
        return [tokenizer.trunk.decode(t) for t in tokens.tolist()]


def get_ocr_metrics(model, tokenizer, image_input, text_input, device_env, max_recursion_length, prompt_token: str):
    metrics = dict()
    with torch.inference_mode():
        m = model.module if hasattr(model, "module") else model
        image_encoding = m.image_encoder(image_input)
        text_input = text_input.clone()
        text_input[text_input == -100] = tokenizer.trunk.pad_token_id
        lengths = (text_input != tokenizer.trunk.pad_token_id).sum(dim=1)
        max_recursion_length = min(max_recursion_length, int(lengths.max().item()))
        preds = generate_ocr(m, tokenizer, image_encoding, device_env, max_recursion_length, prompt_token)
        refs = tokenizer.trunk.batch_decode(text_input)
        preds = [re.sub(r"<.*?>", "", re.sub("\n", " ", t)) for t in preds]
        refs = [re.sub(r"<.*?>", "", re.sub("\n", " ", t)) for t in refs]
        pairs = [(r, p) for r, p in zip(refs, preds) if r and p]
        if not pairs:
            return None, None
        refs, preds = map(list, zip(*pairs))
        preds = [p[0:len(r)] for p, r in zip(preds, refs)]
        metrics["wer"] = wer(refs, preds)
        metrics["cer"] = cer(refs, preds)
        sample = {"image": image_input[0], "original_text": refs[0], "reconstructed_text": preds[0]}
    return metrics, sample
