"""Configuration defaults of the reference (`utils.empty_config`, utils.py:7-57) without fvcore."""
from __future__ import annotations

import yaml


class CfgNode(dict):
    """Attribute-style dict with `merge_from_file`, the subset of fvcore's CfgNode the reference uses."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        self[key] = value

    def merge_from_file(self, path: str) -> None:
        with open(path) as f:
            self.merge_from_dict(yaml.safe_load(f) or {})

    def merge_from_dict(self, src: dict) -> None:
        for k, v in src.items():
            if isinstance(v, dict):
                node = self.get(k)
                if not isinstance(node, CfgNode):
                    node = CfgNode()
                    self[k] = node
                node.merge_from_dict(v)
            else:
                self[k] = v


def empty_config() -> CfgNode:
    c = CfgNode()
    c.merge_from_dict(dict(
        num_epoch=1, num_episodes=1, num_episodes_baseline=1, lr=0.01, clip_value=0.1, entropy_reg_strength=1.0,
        risk_epsilon=0.1, replay_buffer_size=128, replay_buffer_sample_size=32, replay_buffer_score_bound=10,
        loss=dict(BALANCED_ELU_LOSS=False, ELU_LOSS=False),
        summary_name="Try", summary_path="tb_summary", checkpoint_path="checkpoints", reload_checkpoint_path="",
        dataset_path="", val_dataset_path="", instance_path="", sequences_file="", raw_tree_file="", c_best_tree_file="",
        dataset_taxa_list=[], dataset_len_list=[],
        env=dict(batch_size=8, sequence_type="DNA_WITH_GAP"),
        model=dict(vocab_size=4, patch_size=4, fixed_length=1024, embed_dim=32, encoder_attn_layers=2, num_enc_heads=4,
                   num_enc_layers=3),
        ratio_factor=1.0,
    ))
    return c


def inference_config() -> CfgNode:
    """`empty_config()` merged with config/finetune_reinforce_search_example.yaml:24-30 (the shipped inference model)."""
    c = empty_config()
    c.merge_from_dict(dict(num_epoch=1000, num_episodes=10, num_episodes_baseline=50, lr=0.0005, entropy_reg_strength=0.2,
                           risk_epsilon=0.5, env=dict(batch_size=1, sequence_type="DNA_WITH_GAP"),
                           model=dict(vocab_size=4, patch_size=1, embed_dim=64, num_enc_heads=8, num_enc_layers=6)))
    return c
