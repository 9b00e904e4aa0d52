"""Self-supervised heads (reference ``model/self_supervised_learning.py``): protein masked-LM through
the shared ProteinCNN plus an LLM-logit head, and drug SimSiam between GCN and ChemBERTa features.

Same constructor / forward signature / state_dict as the reference ``SSL`` (including the lazily
created SimSiam projectors).  The MLM heads only evaluate the sampled positions:
``F.cross_entropy(..., ignore_index=0)`` ignores every other row, so the logits GEMMs run on the
``B x ceil(0.15*L)`` gathered rows instead of all ``B x L`` -- same loss and gradients, ~7x less work.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import functions as Fn
from . import kernels as K


# ---- mask sampling: same semantics and the same torch-RNG consumption order as the reference's
#      helpers (utils.py:532-554), so a given torch seed yields the reference's mask ------------
def sample_mlm_mask(seq: torch.Tensor, mask_prob: float = 0.15, replace_prob: float = 0.9,
                    ignore_token_ids=(0,), mask_token_id: int = 26):
    """Returns (labels int64 (B,L), masked_seq like seq, sampled positions (B,M) int64 with -1 = unused)."""
    B, Ls = seq.shape
    maskable = torch.ones_like(seq, dtype=torch.bool)
    for t in ignore_token_ids:
        maskable &= seq != t
    max_masked = math.ceil(mask_prob * Ls)
    num_tokens = maskable.sum(dim=-1, keepdim=True)
    excess = (maskable.cumsum(dim=-1) > (num_tokens * mask_prob).ceil())[:, :max_masked]
    rand = torch.rand((B, Ls), device=seq.device).masked_fill(~maskable, -1e9)
    _, idx = rand.topk(max_masked, dim=-1)
    idx = (idx + 1).masked_fill_(excess, 0)                       # 0 = not sampled
    new_mask = torch.zeros((B, Ls + 1), device=seq.device)
    new_mask.scatter_(-1, idx, 1)
    mask = new_mask[:, 1:].bool()
    labels = seq.masked_fill(~mask, 0).long()
    replace = torch.zeros_like(seq).float().uniform_(0, 1) < replace_prob
    masked_seq = seq.clone().detach().masked_fill(mask & replace, mask_token_id)
    return labels, masked_seq, idx - 1


def _sequential(x, seq: nn.Sequential):
    """Run a Linear / BatchNorm1d / ReLU Sequential on the dl_* kernels."""
    for m in seq:
        if isinstance(m, nn.Linear):
            x = Fn.linear(x, m.weight, m.bias)
        elif isinstance(m, nn.BatchNorm1d):
            x = Fn.batch_norm(x, m)
        elif isinstance(m, nn.ReLU):
            x = Fn.ActFn.apply(x, K.ACT_RELU)
        else:  # pragma: no cover
            raise NotImplementedError(type(m))
    return x


def SimSiamMLP(dim, proj_out, hidden_size=512):
    """reference ``self_supervised_learning.py:153-166``."""
    return nn.Sequential(
        nn.Linear(dim, hidden_size, bias=False), nn.BatchNorm1d(hidden_size), nn.ReLU(inplace=True),
        nn.Linear(hidden_size, hidden_size, bias=False), nn.BatchNorm1d(hidden_size), nn.ReLU(inplace=True),
        nn.Linear(hidden_size, proj_out, bias=False), nn.BatchNorm1d(proj_out, affine=False))


def PredictorMLP(dim, proj_out, hidden_size=None):
    """reference ``self_supervised_learning.py:143-151``."""
    hidden_size = dim if hidden_size is None else hidden_size
    return nn.Sequential(nn.Linear(dim, hidden_size), nn.BatchNorm1d(hidden_size), nn.ReLU(inplace=True),
                         nn.Linear(hidden_size, proj_out))


class SimProj(nn.Module):
    """reference ``self_supervised_learning.py:126-141``: the projector is created at the first call."""

    def __init__(self, projection_out, projection_hidden_size=512):
        super().__init__()
        self.projector = None
        self.projection_out = projection_out
        self.projection_hidden_size = projection_hidden_size

    def forward(self, x):
        if self.projector is None:
            self.projector = SimSiamMLP(x.shape[1], self.projection_out, self.projection_hidden_size).to(x.device)
        return _sequential(x, self.projector)


def _neg_cos_loss(x, y):
    """2 - 2 * <l2norm(x), l2norm(y)> per row (reference ``loss_fn`` :184-187)."""
    xn, yn = Fn.L2NormFn.apply(x), Fn.L2NormFn.apply(y)
    return 2 - 2 * (xn.float() * yn.float()).sum(dim=-1)


class SSL(nn.Module):
    def __init__(self, prot_extractor, n_prot_feature, *, drug_ssl_type='simsiam', n_hidden=128, **kwargs):
        super().__init__()
        self.extractor = prot_extractor
        self.to_logits = nn.Linear(128, 26 + 1)
        self.llm_to_logits = nn.Linear(n_prot_feature + 1, 26 + 1)
        self.n_prot_feature = n_prot_feature
        self.drug_ssl_type = drug_ssl_type
        self.net = SimProj(n_hidden)
        self.llm_net = SimProj(n_hidden)
        if self.drug_ssl_type == 'simsiam':
            self.predictor = PredictorMLP(n_hidden, n_hidden, n_hidden * 4)
        else:
            self.temperature = 0.1

    # ---------------------------------------------------------------- drug branch
    def drug_simsiam(self, vd, xd):
        one, two = vd.reshape(-1, vd.shape[-1]), xd.reshape(-1, xd.shape[-1])
        proj_one, proj_two = self.net(one), self.llm_net(two)
        pred_one = _sequential(proj_one, self.predictor)
        pred_two = _sequential(proj_two, self.predictor)
        with torch.no_grad():       # second pass like the reference (BN running stats move twice)
            target_one, target_two = self.net(one.detach()), self.llm_net(two.detach())
        return (_neg_cos_loss(pred_one, target_two) + _neg_cos_loss(pred_two, target_one)).mean()

    def drug_simclr(self, vd, xd):
        """NT-Xent (reference :35-41, :168-182; unreachable with the default drug_ssl_type)."""
        q = self.net(vd.reshape(-1, vd.shape[-1]))
        k = self.llm_net(xd.reshape(-1, xd.shape[-1]))
        b = q.shape[0]
        n = 2 * b
        projs = torch.cat((q, k))
        logits = Fn.MatmulNTFn.apply(projs, projs)
        eye = torch.eye(n, device=logits.device, dtype=torch.bool)
        logits = logits[~eye].reshape(n, n - 1) / self.temperature
        labels = torch.cat((torch.arange(b, device=logits.device) + b - 1, torch.arange(b, device=logits.device)))
        return Fn.CrossEntropyFn.apply(logits, labels, -100)

    # ---------------------------------------------------------------- protein branch
    def prot_mlm(self, seq, extractor, xp, fill_bit, mode, mask_ignore_token_ids=(0,), mask_prob=0.15,
                 replace_prob=0.9, pad_token_id=0, mask_token_id=26):
        if getattr(self, "_mask_override", None) is not None:      # tests: a mask sampled elsewhere
            labels, masked_seq, pos = (t.to(seq.device) for t in self._mask_override)
        else:
            labels, masked_seq, pos = sample_mlm_mask(seq, mask_prob, replace_prob, mask_ignore_token_ids,
                                                      mask_token_id)
        B, M = pos.shape
        valid = pos >= 0
        posc = pos.clamp(min=0)
        lab_g = torch.where(valid, labels.gather(1, posc), torch.zeros_like(posc)).reshape(-1)
        losses = []
        if mode != 'xp':
            emb = extractor(masked_seq, fill_bit)                                   # (B, L, 128)
            emb_g = emb.gather(1, posc.unsqueeze(-1).expand(B, M, emb.shape[-1]))
            logits = Fn.linear(emb_g, self.to_logits.weight, self.to_logits.bias)
            losses.append(Fn.CrossEntropyFn.apply(logits, lab_g, pad_token_id))
        if mode != 'vp':
            W, b = self.llm_to_logits.weight, self.llm_to_logits.bias
            xg = xp.gather(1, posc.unsqueeze(-1).expand(B, M, xp.shape[-1]))
            if xp.shape[-1] == self.n_prot_feature + 1:                            # reference form: cat(xp, bit)
                llm_logits = Fn.linear(xg, W, b)
            else:                                                                   # lazy form: raw xp + fill bit
                llm_logits = Fn.linear(xg, W[:, :-1], b).float()
                llm_logits = llm_logits + fill_bit.gather(1, posc).unsqueeze(-1) * W[:, -1]
            losses.append(Fn.CrossEntropyFn.apply(llm_logits, lab_g, pad_token_id))
        return sum(losses) / len(losses)

    def forward(self, vp, xp, fill_bit_p, vd, xd, p_mode='double'):
        prot_ssl_loss = self.prot_mlm(vp, self.extractor, xp, fill_bit_p, p_mode)
        if (vd is None) or (xd is None):
            drug_ssl_loss = 0
        elif self.drug_ssl_type == 'simsiam':
            drug_ssl_loss = self.drug_simsiam(vd, xd)
        else:
            drug_ssl_loss = self.drug_simclr(vd, xd)
        return {'prot_ssl': prot_ssl_loss, 'drug_ssl': drug_ssl_loss}
