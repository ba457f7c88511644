"""CPU probe: logit error of BERT-base (random init, like BASELINE configs[3]) when every Linear / attention matmul
rounds its operands as a given tensor-core scheme would.  Decides which precision modes can meet the 1e-3 bar.
Checker-side tooling only (uses the oracle restatement's op sequence)."""
import math, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, "tests")
from conftest import load_golden, rel_err
import json

def rnd(x, kind):
    if kind == "fp32": return x
    if kind == "bf16": return x.to(torch.bfloat16).to(x.dtype)
    if kind == "fp16": return x.to(torch.float16).to(x.dtype)
    if kind == "tf32":
        i = x.float().view(torch.int32); i = (i + 0x1000) & ~0x1FFF; return i.view(torch.float32).to(x.dtype)
    if kind == "fp16x2":  # hi + lo, both fp16 (lo may be subnormal)
        hi = x.to(torch.float16).to(x.dtype); lo = (x - hi).to(torch.float16).to(x.dtype); return hi + lo
    raise ValueError(kind)

def logits(state, ids, mask, seg, nh, a_kind, w_kind, att_kind, dt=torch.float32, eps=1e-12):
    st = {k: v.to(dt) if v.dtype.is_floating_point else v for k, v in state.items()}
    g = lambda k: st[k]
    lin = lambda x, w, b: F.linear(rnd(x, a_kind), rnd(g(w), w_kind), g(b))
    N, L = ids.shape
    x = F.embedding(ids, g("bert.embeddings.word_embeddings.weight")) + F.embedding(seg, g("bert.embeddings.token_type_embeddings.weight")) + g("bert.embeddings.position_embeddings.weight")[:L][None]
    H = x.shape[-1]; dh = H // nh
    x = F.layer_norm(x, (H,), g("bert.embeddings.LayerNorm.weight"), g("bert.embeddings.LayerNorm.bias"), eps)
    kb = torch.zeros(N, 1, 1, L, dtype=dt); kb.masked_fill_(mask[:, None, None, :] == 0, -1e30)
    nl = 1 + max(int(k.split(".")[3]) for k in state if k.startswith("bert.encoder.layer."))
    for i in range(nl):
        p = f"bert.encoder.layer.{i}."
        sp = lambda t: t.reshape(N, L, nh, dh).transpose(1, 2)
        q = sp(lin(x, p + "attention.self.query.weight", p + "attention.self.query.bias"))
        k = sp(lin(x, p + "attention.self.key.weight", p + "attention.self.key.bias"))
        v = sp(lin(x, p + "attention.self.value.weight", p + "attention.self.value.bias"))
        ak, bk = att_kind if isinstance(att_kind, tuple) else (att_kind, att_kind)
        s = rnd(q, ak) @ rnd(k, bk).transpose(-1, -2) / math.sqrt(dh) + kb
        pr = torch.softmax(s, -1)
        ctx = (rnd(pr, ak) @ rnd(v, bk)).transpose(1, 2).reshape(N, L, H)
        y = lin(ctx, p + "attention.output.dense.weight", p + "attention.output.dense.bias")
        x = F.layer_norm(x + y, (H,), g(p + "attention.output.LayerNorm.weight"), g(p + "attention.output.LayerNorm.bias"), eps)
        y = F.gelu(lin(x, p + "intermediate.dense.weight", p + "intermediate.dense.bias"))
        y = lin(y, p + "output.dense.weight", p + "output.dense.bias")
        x = F.layer_norm(x + y, (H,), g(p + "output.LayerNorm.weight"), g(p + "output.LayerNorm.bias"), eps)
    pooled = torch.tanh(F.linear(x[:, 0], g("bert.pooler.dense.weight"), g("bert.pooler.dense.bias")))
    return F.linear(pooled, g("classifier.weight"), g("classifier.bias"))

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "mid"
    from transformers import BertConfig, BertForSequenceClassification
    g = load_golden(f"bert_{name}"); cfg = json.loads(str(g["config_json"]))
    N, P, L, _ = (int(x) for x in g["shape"])
    torch.manual_seed(int(g["weight_seed"]))
    keep = ("hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size", "vocab_size", "max_position_embeddings", "type_vocab_size", "initializer_range", "layer_norm_eps", "hidden_act")
    m = BertForSequenceClassification(BertConfig(**{k: cfg[k] for k in keep if k in cfg}, hidden_dropout_prob=0.1)).eval()
    state = m.state_dict(); nh = cfg["num_attention_heads"]
    f = lambda k: torch.from_numpy(g[k].astype(np.int64)).reshape(N * P, L)
    ids, mask, seg = f("pos_bert_input"), f("pos_mask"), f("pos_seg")
    nmax = int(sys.argv[2]) if len(sys.argv) > 2 else N * P
    ids, mask, seg = ids[:nmax], mask[:nmax], seg[:nmax]
    with torch.no_grad():
        ref64 = logits(state, ids, mask, seg, nh, "fp32", "fp32", "fp32", dt=torch.float64).numpy()
        ref32 = logits(state, ids, mask, seg, nh, "fp32", "fp32", "fp32").numpy()
        gold = g["logits"][:nmax]
        print("shape", ids.shape, "logit scale", np.abs(gold).max(), "fp32 vs golden", rel_err(ref32[:, 1], gold[:, 1], 1e-2), "fp32 vs fp64", rel_err(ref32[:, 1], ref64[:, 1], 1e-2))
        for a_kind, w_kind, att_kind in [("bf16", "bf16", "bf16"), ("tf32", "tf32", "tf32"), ("fp16", "fp16", "fp16"), ("fp16", "fp32", "fp16"), ("fp16", "fp32", "fp32"), ("fp32", "fp16", "fp32"), ("fp32", "fp32", "fp16"), ("fp16x2", "fp16", "fp16x2"), ("fp16", "fp16x2", "fp16"), ("fp16", "fp32", ("fp16", "fp32")), ("fp32", "fp16", ("fp32", "fp16"))]:
            out = logits(state, ids, mask, seg, nh, a_kind, w_kind, att_kind).numpy()
            print(f"A={a_kind:7s} W={w_kind:7s} att={str(att_kind):20s}: score err vs golden(floor 1e-2) {rel_err(out[:, 1], gold[:, 1], 1e-2):.2e}   logits (5% floor) {rel_err(out, gold, 0.05 * float(np.abs(gold).max())):.2e}")
