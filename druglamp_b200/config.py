"""Config objects with the keys the hot path reads (reference ``configs/default_config.py``).

The reference builds yacs ``CfgNode``s; only attribute/item access and ``clone()`` are used on the
model side, so a small dict subclass is enough and keeps yacs optional.  A real yacs node passed by
``main.py`` works as well (it is a dict subclass with the same access patterns).
"""
from __future__ import annotations


class CfgNode(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        out = CfgNode()
        for k, v in self.items():
            out[k] = v.clone() if isinstance(v, CfgNode) else (list(v) if isinstance(v, list) else v)
        return out


def get_cfg_defaults() -> CfgNode:
    """The sub-trees ``DrugLAMPBase.__init__`` reads (``configs/default_config.py:4-61``,
    ``model/basic_model.py:60-69,93-94``)."""
    c = CfgNode()
    c.DRUG = CfgNode(NODE_IN_FEATS=75, PADDING=True, MAX_NODES=512)
    c.PROTEIN = CfgNode(KERNEL_SIZE=[3, 6, 9], PADDING=True, SEQ_LEN=9 * 256, SITE_LEN=9)
    c.DECODER = CfgNode(NAME="MLP", IN_DIM=256, HIDDEN_DIM=512, OUT_DIM=128, BINARY=1)
    c.RS = CfgNode(MAX_MARGIN=0.5, RESET_EPOCH=100)
    return c


def get_model_defaults(hidden_size: int) -> CfgNode:
    """PMMA hyper-config (``configs/default_config.py:67-89``): width 2*hidden, 4 heads, 4 layers,
    attention dropout 0, dropout 0.1, MHLA dropout 0, 256 protein sites and 256 "mol" tokens."""
    config = CfgNode()
    config.n_output = 1
    config.hidden_size = hidden_size * 2
    config.num_features_llm = config.hidden_size
    config.mlha_dropout = 0
    config.transformer = CfgNode(num_heads=4, num_p_plus_s_layers=4, attention_dropout_rate=0,
                                 dropout_rate=0.1)
    config.classifier = "token"
    config.representation_size = None
    config.feat_len = 256
    config.mol_len = config.feat_len
    return config
