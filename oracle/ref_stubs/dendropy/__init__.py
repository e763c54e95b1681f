"""Empty stub: dendropy is imported but unused on the Argmax path (finetune_rl_search.py:16)."""
