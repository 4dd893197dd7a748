"""TEST INFRASTRUCTURE (oracle): closed-form backward of one GAT layer's gather-attend-aggregate step over a batch of STAR egonets
(data_loader/dataset.py:404-437), i.e. the autograd of model_zoo.py:84-96,106-114 written out per egonet - the algorithm a
star-specialised backward kernel has to implement (kernel work queue item 1).  Checked against torch autograd
of oracle.taxo_oracle.gat_layer in tests/test_oracle_golden.py; nothing under taxoexpan_b200/ imports this.

Inputs, all float64 numpy unless noted, per batch:
  n_gp, n_sib [G] ints;  ft [N, H, D] projected features;  g [N, H, D] = d(loss)/d(aggregate output);  attn_l, attn_r [H, D];
  keepw [E, H] = attention-dropout factor per EDGE ID (keep / (1 - p), all ones without dropout);  neg_slope.
Edge ids / node order as tx_star_batch_structure: egonet with first node o, first edge q, a = n_gp, s = n_sib:
  nodes  o .. o+a-1 grand-parents, o+a anchor, o+a+1 .. siblings;   edges  q+k: gp_k -> anchor,  q+a+k: anchor -> sib_k,
  q+a+s+t: self loop of local node t.
Returns d(ft) [N, H, D], d(attn_l) [H, D], d(attn_r) [H, D].

Structure that makes the kernel cheap:
  * grand-parent k has ONE in-edge (its self loop): alpha = 1 and the softmax backward vanishes -> ds_self = 0, da2_k = 0;
    d(ft_k) = kw_self g_k + alpha~(k->anchor) g_anchor + ds(k->anchor) attn_l.
  * sibling k has TWO in-edges {anchor, self}: closed-form 2x2 softmax backward from the dots <g_s, ft_anchor>, <g_s, ft_s>;
    d(ft_s) = alpha~_self g_s + ds_self attn_l + (ds_anchor + ds_self) attn_r, and it feeds the anchor through
    alpha~(anchor->s) g_s and ds_anchor - a LINEAR contribution, so sibling chunks can be reduced in any fixed order.
  * anchor: in-edges {gp_0.., self}: dots <g_anchor, ft_gp_k>, <g_anchor, ft_anchor>, softmax backward over a + 1 logits.
"""
import numpy as np


def _lrelu(x, slope):
    return np.where(x > 0, x, slope * x)


def star_gat_backward(n_gp, n_sib, ft, g, attn_l, attn_r, keepw, neg_slope=0.2):
    N, H, D = ft.shape
    dft = np.zeros_like(ft)
    dal = np.zeros_like(attn_l)
    dar = np.zeros_like(attn_r)
    a1 = (ft * attn_l[None]).sum(-1)          # [N, H]
    a2 = (ft * attn_r[None]).sum(-1)
    o = q = 0
    for a, s in zip(np.asarray(n_gp).tolist(), np.asarray(n_sib).tolist()):
        n = a + 1 + s
        self0 = q + a + s
        A = o + a                              # anchor node
        for h in range(H):
            # ---- anchor: softmax over {gp_0 .. gp_{a-1}, self} ----
            src = list(range(o, o + a)) + [A]
            eid = [q + k for k in range(a)] + [self0 + a]
            logit = np.array([a1[j, h] + a2[A, h] for j in src])
            e = _lrelu(logit, neg_slope)
            al = np.exp(e - e.max()); al /= al.sum()
            kw = np.array([keepw[x, h] for x in eid])
            dd = np.array([g[A, h] @ ft[j, h] for j in src])              # d(alpha~) per in-edge
            dalpha = dd * kw
            de = al * (dalpha - (al * dalpha).sum())
            ds = de * np.where(e > 0, 1.0, neg_slope)
            da2_A = ds.sum()
            da1 = np.zeros(n)                   # per local node: sum of ds over its OUT-edges
            da2 = np.zeros(n)
            da2[a] = da2_A
            for k in range(a):
                da1[k] += ds[k]                 # gp_k -> anchor
            da1[a] += ds[a]                     # anchor self loop
            acc_A = al[a] * kw[a] * g[A, h]     # alpha~(self) g_anchor
            for k in range(a):
                # grand-parent k: own output = kw_self ft_k (alpha = 1): d(ft_k) gets kw_self g_k; ds_self = 0
                gk = o + k
                dft[gk, h] += keepw[self0 + k, h] * g[gk, h] + al[k] * kw[k] * g[A, h]
            # ---- siblings: 2-way softmax {anchor -> s, self} ----
            for k in range(s):
                S = o + a + 1 + k
                t = a + 1 + k
                e1 = _lrelu(a1[A, h] + a2[S, h], neg_slope)
                e2 = _lrelu(a1[S, h] + a2[S, h], neg_slope)
                m = max(e1, e2)
                x1, x2 = np.exp(e1 - m), np.exp(e2 - m)
                al1, al2 = x1 / (x1 + x2), x2 / (x1 + x2)
                kw1, kw2 = keepw[q + a + k, h], keepw[self0 + t, h]
                d1 = (g[S, h] @ ft[A, h]) * kw1
                d2 = (g[S, h] @ ft[S, h]) * kw2
                tsum = al1 * d1 + al2 * d2
                ds1 = al1 * (d1 - tsum) * (1.0 if e1 > 0 else neg_slope)
                ds2 = al2 * (d2 - tsum) * (1.0 if e2 > 0 else neg_slope)
                da2[t] = ds1 + ds2
                da1[t] = ds2                    # only out-edge of a sibling: its self loop
                da1[a] += ds1                   # anchor -> sibling
                acc_A = acc_A + al1 * kw1 * g[S, h]
                dft[S, h] += al2 * kw2 * g[S, h]
            dft[A, h] += acc_A
            # ---- half-logit terms and d(attn) ----
            for t in range(n):
                j = o + t
                dft[j, h] += da1[t] * attn_l[h] + da2[t] * attn_r[h]
                dal[h] += da1[t] * ft[j, h]
                dar[h] += da2[t] * ft[j, h]
        o += n
        q += 2 * n - 1
    return dft, dal, dar
