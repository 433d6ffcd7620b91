"""Synthetic workload generator for the Cruller train step (BASELINE.json configs).

Pure CPU torch; shared by the benchmark, the tests and the oracle so that every arm sees byte-identical
inputs for a given seed. Mirrors what the reference's data pipeline hands to ``train_step``
(/root/reference/src/pixparse/task/task_cruller_pretrain.py:236-242, data/preprocess.py:62-68,97-101):

    image  (B, 1, H, W) float32, already normalised with mean 0.5 / std 0.5
    text   (B, Lt) int64   = [<s_pretrain>, tok..., </s>, <pad>...]
    target (B, Lt) int64   = text with PAD -> -100 and the prompt token -> -100
"""
import torch

PAD_ID = 1
EOS_ID = 2
BART_VOCAB = 50265
# tokens are added to the tokenizer in sorted() order: '<s_pretrain>' < '<sep/>'
# (task_cruller_pretrain.py:92-99) -> ids 50265 and 50266, vocab 50267
S_PRETRAIN_ID = 50265
SEP_ID = 50266
PRETRAIN_VOCAB = 50267


def synthetic_pages_u8(batch, height=1100, width=850, seed=0):
    """uint8 'L' pages: white background with random dark rectangles ("text lines")."""
    g = torch.Generator().manual_seed(seed)
    pages = torch.full((batch, height, width), 255, dtype=torch.uint8)
    for b in range(batch):
        n = int(torch.randint(20, 60, (1,), generator=g))
        ys = torch.randint(0, height - 12, (n,), generator=g)
        xs = torch.randint(0, width - 40, (n,), generator=g)
        hs = torch.randint(4, 12, (n,), generator=g)
        ws = torch.randint(20, width // 2, (n,), generator=g)
        vs = torch.randint(0, 96, (n,), generator=g)
        for y, x, h, w, v in zip(ys.tolist(), xs.tolist(), hs.tolist(), ws.tolist(), vs.tolist()):
            pages[b, y:y + h, x:min(width, x + w)] = v
    return pages


def synthetic_batch(batch, image_size=(576, 448), text_len=513, vocab=PRETRAIN_VOCAB, seed=0,
                    min_eos_frac=0.75, start_id=S_PRETRAIN_ID):
    """Return (image, text, target) exactly in the layout the reference's train_step consumes."""
    g = torch.Generator().manual_seed(seed)
    H, W = image_size
    image = torch.rand((batch, 1, H, W), generator=g, dtype=torch.float32)
    image = (image - 0.5) / 0.5
    text = torch.randint(3, min(vocab, BART_VOCAB), (batch, text_len), generator=g, dtype=torch.int64)
    text[:, 0] = start_id
    lo = max(2, int(min_eos_frac * text_len))
    eos_pos = torch.randint(lo, text_len, (batch,), generator=g)
    for b in range(batch):
        p = int(eos_pos[b])
        text[b, p] = EOS_ID
        text[b, p + 1:] = PAD_ID
    target = text.clone()
    target[text == PAD_ID] = -100
    target[:, 0] = -100   # prompt_end_token == task_start_token at position 0 (data/preprocess.py:97-101)
    return image, text, target


class SyntheticBartTokenizer:
    """Tokenizer with facebook/bart-large's id layout (pad 1, eos 2, bos 0, 50265 entries) for offline runs: it can
    add and look up special tokens, which is all the train step needs when the token ids are synthetic."""

    def __init__(self):
        self.pad_token_id, self.eos_token_id, self.bos_token_id, self.unk_token_id = PAD_ID, EOS_ID, 0, 3
        self.pad_token, self.eos_token, self.bos_token = "<pad>", "</s>", "<s>"
        self._size = BART_VOCAB
        self._added = {}

    def __len__(self):
        return self._size

    def add_special_tokens(self, d):
        n = 0
        for t in d.get("additional_special_tokens", []):
            if t not in self._added:
                self._added[t] = self._size
                self._size += 1
                n += 1
        return n

    def add_tokens(self, toks):
        return self.add_special_tokens({"additional_special_tokens": list(toks)})

    def convert_tokens_to_ids(self, t):
        if isinstance(t, (list, tuple)):
            return [self.convert_tokens_to_ids(x) for x in t]
        base = {"<s>": 0, "<pad>": 1, "</s>": 2, "<unk>": 3}
        return self._added.get(t, base.get(t, 3))

    def encode(self, text, add_special_tokens=False):
        return [self.convert_tokens_to_ids(text)]

    def decode(self, ids, **kw):
        inv = {v: k for k, v in self._added.items()}
        return " ".join(inv.get(int(i), f"<{int(i)}>") for i in ids)

    def batch_decode(self, batch, **kw):
        return [self.decode(x) for x in batch]


class CharTokenizer:
    """Deterministic character-level stand-in for a HF tokenizer's __call__ interface (offline tests of the annotation
    preprocessors): special tokens registered with add_special_tokens map to single ids, every other character to
    4 + ord(c) % 5000; supports max_length / padding='max_length' / truncation like the calls in data/preprocess.py."""

    class _Out:
        def __init__(self, ids):
            self.input_ids = ids

    def __init__(self):
        self.pad_token_id, self.eos_token_id, self.bos_token_id = PAD_ID, EOS_ID, 0
        self.pad_token, self.eos_token, self.bos_token = "<pad>", "</s>", "<s>"
        self._special = {"<s>": 0, "<pad>": 1, "</s>": 2, "<unk>": 3}
        self._next = 6000

    def __len__(self):
        return self._next

    def add_special_tokens(self, d):
        n = 0
        for t in d.get("additional_special_tokens", []):
            if t not in self._special:
                self._special[t] = self._next
                self._next += 1
                n += 1
        return n

    def convert_tokens_to_ids(self, t):
        return self._special.get(t, 3)

    def _encode(self, text):
        ids, i = [], 0
        specials = sorted(self._special, key=len, reverse=True)
        while i < len(text):
            for s in specials:
                if text.startswith(s, i):
                    ids.append(self._special[s])
                    i += len(s)
                    break
            else:
                ids.append(4 + ord(text[i]) % 5000)
                i += 1
        return ids

    def __call__(self, text, add_special_tokens=False, return_tensors='pt', max_length=None, padding=None,
                 truncation=False):
        ids = self._encode(text)
        if truncation and max_length is not None:
            ids = ids[:max_length]
        if padding == 'max_length' and max_length is not None:
            ids = ids + [self.pad_token_id] * (max_length - len(ids))
        return CharTokenizer._Out(torch.tensor([ids], dtype=torch.int64))


def synthetic_ocr_annotation(seed, max_pages=4):
    """A pixparse OCR annotation: {'pages': [{'text': [line, ...]}, ...]} with some empty pages (data/preprocess.py:43)."""
    import random
    r = random.Random(seed)
    pages = []
    for _ in range(r.randint(1, max_pages)):
        if r.random() < 0.3:
            pages.append({"text": []})
        else:
            pages.append({"text": ["".join(r.choice("abcdefghij klmnop") for _ in range(r.randint(3, 30)))
                                   for _ in range(r.randint(1, 5))]})
    if not any(p["text"] for p in pages):
        pages[r.randrange(len(pages))]["text"] = ["fallback line"]
    return {"pages": pages}
