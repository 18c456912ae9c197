"""Synthetic private shards (there is no dataset access on the target box)."""
from .synthetic import (LINEAR_TRUTH, ShardSpec, dirichlet_label_shards, image_shard, iid_label_shards,
                        label_skew_shards, linear_regression_shard, token_shard)

__all__ = ["LINEAR_TRUTH", "ShardSpec", "linear_regression_shard", "image_shard", "token_shard",
           "iid_label_shards", "label_skew_shards", "dirichlet_label_shards"]
