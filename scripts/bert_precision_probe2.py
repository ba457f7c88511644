"""CPU probe (round 2): which of the 3 bf16 products of the parity mode could be dropped per GEMM.  Baseline = every operand as
hi + lo bf16 (what bf16x3 feeds the tensor core); a variant rounds ONE operand of ONE GEMM family to plain bf16 (= dropping that
operand's lo product there).  Reports the score error against the reference golden on BERT-base sequences.
    python scripts/bert_precision_probe2.py [golden name] [n sequences]"""
import json, math, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, "tests")
from conftest import load_golden, rel_err

def r2(x):  # hi + lo, both bf16
    hi = x.to(torch.bfloat16).to(x.dtype)
    return hi + (x - hi).to(torch.bfloat16).to(x.dtype)
def r1(x):
    return x.to(torch.bfloat16).to(x.dtype)

def logits(state, ids, mask, seg, nh, drop, eps=1e-12):
    """drop: set of (gemm, side) with gemm in qkv/out/ffn1/ffn2/qk/pv and side in a/b whose lo product is dropped."""
    g = lambda k: state[k]
    ra = lambda x, t: r1(x) if (t, "a") in drop else r2(x)
    rb = lambda x, t: r1(x) if (t, "b") in drop else r2(x)
    lin = lambda x, w, b, t: F.linear(ra(x, t), rb(g(w), t), g(b))
    N, L = ids.shape
    x = F.embedding(ids, g("bert.embeddings.word_embeddings.weight")) + F.embedding(seg, g("bert.embeddings.token_type_embeddings.weight")) + g("bert.embeddings.position_embeddings.weight")[:L][None]
    H = x.shape[-1]; dh = H // nh
    x = F.layer_norm(x, (H,), g("bert.embeddings.LayerNorm.weight"), g("bert.embeddings.LayerNorm.bias"), eps)
    kb = torch.zeros(N, 1, 1, L); kb.masked_fill_(mask[:, None, None, :] == 0, -1e30)
    nl = 1 + max(int(k.split(".")[3]) for k in state if k.startswith("bert.encoder.layer."))
    for i in range(nl):
        p = f"bert.encoder.layer.{i}."
        sp = lambda t: t.reshape(N, L, nh, dh).transpose(1, 2)
        q = sp(lin(x, p + "attention.self.query.weight", p + "attention.self.query.bias", "qkv"))
        k = sp(lin(x, p + "attention.self.key.weight", p + "attention.self.key.bias", "qkv"))
        v = sp(lin(x, p + "attention.self.value.weight", p + "attention.self.value.bias", "qkv"))
        s = ra(q, "qk") @ rb(k, "qk").transpose(-1, -2) / math.sqrt(dh) + kb
        pr = torch.softmax(s, -1)
        ctx = (ra(pr, "pv") @ rb(v, "pv")).transpose(1, 2).reshape(N, L, H)
        y = lin(ctx, p + "attention.output.dense.weight", p + "attention.output.dense.bias", "out")
        x = F.layer_norm(x + y, (H,), g(p + "attention.output.LayerNorm.weight"), g(p + "attention.output.LayerNorm.bias"), eps)
        y = F.gelu(lin(x, p + "intermediate.dense.weight", p + "intermediate.dense.bias", "ffn1"))
        y = lin(y, p + "output.dense.weight", p + "output.dense.bias", "ffn2")
        x = F.layer_norm(x + y, (H,), g(p + "output.LayerNorm.weight"), g(p + "output.LayerNorm.bias"), eps)
    pooled = torch.tanh(F.linear(x[:, 0], g("bert.pooler.dense.weight"), g("bert.pooler.dense.bias")))
    return F.linear(pooled, g("classifier.weight"), g("classifier.bias"))

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "base"
    from transformers import BertConfig, BertForSequenceClassification
    g = load_golden(f"bert_{name}"); cfg = json.loads(str(g["config_json"]))
    N, P, L, _ = (int(x) for x in g["shape"])
    torch.manual_seed(int(g["weight_seed"]))
    keep = ("hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size", "vocab_size", "max_position_embeddings", "type_vocab_size", "initializer_range", "layer_norm_eps", "hidden_act")
    m = BertForSequenceClassification(BertConfig(**{k: cfg[k] for k in keep if k in cfg}, hidden_dropout_prob=0.1)).eval()
    state = m.state_dict(); nh = cfg["num_attention_heads"]
    f = lambda k: torch.from_numpy(g[k].astype(np.int64)).reshape(N * P, L)
    nmax = int(sys.argv[2]) if len(sys.argv) > 2 else N * P
    ids, mask, seg = f("pos_bert_input")[:nmax], f("pos_mask")[:nmax], f("pos_seg")[:nmax]
    gold = g["logits"][:nmax]
    variants = [set()] + [{(t, s)} for t in ("qkv", "out", "ffn1", "ffn2", "qk", "pv") for s in ("a", "b")]
    variants += [{("ffn1", "a"), ("ffn2", "a")}, {("ffn1", "b"), ("ffn2", "b")}, {("pv", "a"), ("qk", "a")}, {("ffn1", "a"), ("ffn2", "a"), ("pv", "a")}]
    with torch.no_grad():
        for d in variants:
            out = logits(state, ids, mask, seg, nh, d).numpy()
            e = np.abs(out[:, 1] - gold[:, 1]) / np.maximum(np.abs(gold[:, 1]), 1e-2)
            print(f"drop {sorted(d)!s:60s} score err max {e.max():.2e} median {np.median(e):.2e}", flush=True)
