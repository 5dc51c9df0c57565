"""YAML -> model kwargs, mirroring the reference's yacs defaults without yacs.

Reference: utils/config.py:9-61 (defaults), :76-93 (YAML merged over defaults, unknown keys accepted),
:96-117 (get_model_config: the whitelist of keys forwarded to MCA(**kwargs)).
"""
from __future__ import annotations

import copy
from typing import Any, Dict

import yaml

# utils/config.py:9-61 — only the keys the hot path or its launch shell read
TRAIN_DEFAULTS: Dict[str, Any] = {
    "encoder_configs": {},
    "modality_config": {},
    "restart": "",
    "epochs": 3,
    "start_epoch": 0,
    "batch_size": 32,
    "n_step_checkpoint": 0,
    "num_warmup_steps": 3000,
    "lr_scheduler_type": "cosine",
    "lr": 1e-4,
    "clip": 0.0,
    "hidden_size": 512,
    "layers": 10,
    "heads": 8,
    "dim_head": 64,
    "ff_mult": 4,
    "num_fusion_tokens": 256,
    "seed": 42,
    "mean_pool": False,
    "dropout": 0.1,
    "zorro": False,
    "eao": False,
    "bimodal_contrastive": True,
    "non_fusion_fcl": True,
    "fcl": True,
    "no_fusion": False,
    "fcl_root": [1, 2, 3, 4],
    "fusion_combos": [4, 3, 2],
    "return_logits": True,
    "predrop": False,
}


class Config(dict):
    """dict with attribute access (stands in for yacs CfgNode with new_allowed=True)."""

    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def training_config(path_or_dict) -> Config:
    """Defaults overlaid with the YAML file (utils/config.py:76-93) — no output-dir side effects."""
    cfg = Config(copy.deepcopy(TRAIN_DEFAULTS))
    if isinstance(path_or_dict, dict):
        loaded = path_or_dict
    else:
        with open(path_or_dict, "r") as f:
            loaded = yaml.safe_load(f)
    for k, v in (loaded or {}).items():
        cfg[k] = v
    return cfg


def get_model_config(config) -> Dict[str, Any]:
    """utils/config.py:96-117: the kwargs handed to MCA(**model_config)."""
    return {
        "dim": config["hidden_size"],
        "depth": config["layers"],
        "heads": config["heads"],
        "dim_head": config["dim_head"],
        "ff_mult": config["ff_mult"],
        "num_fusion_tokens": config["num_fusion_tokens"],
        "encoder_configs": config["encoder_configs"],
        "batch_size": config["batch_size"],
        "fcl": config["fcl"],
        "fcl_root": config["fcl_root"],
        "bimodal_contrastive": config["bimodal_contrastive"],
        "non_fusion_fcl": config["non_fusion_fcl"],
        "fusion_combos": config["fusion_combos"],
        "zorro": config["zorro"],
        "eao": config["eao"],
        "no_fusion": config["no_fusion"],
        "mean_pool": config["mean_pool"],
    }


# ---------------------------------------------------------------------------------------------------------------
# The five BASELINE.json configurations, as YAML-equivalent dicts (shapes from configs/CMU_config1.yaml:1-33,
# configs/CMU_config1_z.yaml:1-29, configs/CMU_config1_d40.yaml, configs/TCGA_config1.yaml:1-35).
def _cmu_encoders():
    return {
        "COVAREP": {"type": "EmbeddedSequenceEncoder", "input_size": 74, "max_tokens": 1500},
        "FACET": {"type": "EmbeddedSequenceEncoder", "input_size": 35, "max_tokens": 450},
        "OpenFace": {"type": "EmbeddedSequenceEncoder", "input_size": 713, "max_tokens": 450},
        "glove_vectors": {"type": "EmbeddedSequenceEncoder", "input_size": 300, "max_tokens": 50},
    }


def _cmu_modalities(dropout=None):
    out = {}
    for name, (emb, pad) in {"COVAREP": (74, 1500), "FACET": (35, 450), "OpenFace": (713, 450),
                             "glove_vectors": (300, 50)}.items():
        out[name] = {"type": "embedded_sequence", "pad_len": pad, "data_col_name": "data", "pad_token": -10000,
                     "embedding_size": emb}
        if dropout is not None:
            out[name]["dropout"] = dropout
    return out


_COMMON = {"num_fusion_tokens": 88, "batch_size": 8, "seed": 43, "lr": 1e-4, "layers": 5, "clip": 2.0,
           "fcl_root": [0, 1, 2, 3], "fusion_combos": [4, 3, 2]}

NAMED_CONFIGS: Dict[str, Dict[str, Any]] = {
    "CMU_config1": dict(_COMMON, encoder_configs=_cmu_encoders(), modality_config=_cmu_modalities(),
                        bimodal_contrastive=False, non_fusion_fcl=False, fcl=True, zorro=False),
    "CMU_config1_d40": dict(_COMMON, encoder_configs=_cmu_encoders(), modality_config=_cmu_modalities(0.4),
                            predrop=True, bimodal_contrastive=False, non_fusion_fcl=False, fcl=True, zorro=False),
    "CMU_config1_z": dict(_COMMON, encoder_configs=_cmu_encoders(), modality_config=_cmu_modalities(),
                          bimodal_contrastive=False, non_fusion_fcl=False, fcl=False, zorro=True),
    "TCGA_config1": dict(
        _COMMON,
        encoder_configs={
            "gene": {"type": "TabularEncoder", "num_embeddings": 800, "max_tokens": 800, "max_value": 100},
            "protein": {"type": "TabularEncoder", "num_embeddings": 198, "max_tokens": 198, "max_value": 100},
            "methylation": {"type": "TabularEncoder", "num_embeddings": 800, "max_tokens": 800, "max_value": 100},
            "mirna": {"type": "TabularEncoder", "num_embeddings": 662, "max_tokens": 662, "max_value": 100},
        },
        modality_config={
            "gene": {"type": "sequence", "pad_len": 800, "data_col_name": "values", "pad_token": -10000},
            "protein": {"type": "sequence", "pad_len": 198, "data_col_name": "values", "pad_token": -10000},
            "methylation": {"type": "sequence", "pad_len": 800, "data_col_name": "values", "pad_token": -10000},
            "mirna": {"type": "sequence", "pad_len": 662, "data_col_name": "values", "pad_token": -10000},
        },
        bimodal_contrastive=True, non_fusion_fcl=True, fcl=True, zorro=False),
}
# the "everything at once" baseline (configs/CMU_config1_EAO.yaml:1-28): pair passes, mean pooling, no fusion tokens
NAMED_CONFIGS["CMU_config1_EAO"] = dict(NAMED_CONFIGS["CMU_config1"], bimodal_contrastive=True, non_fusion_fcl=True,
                                        fcl=True, fcl_root=[0, 1], fusion_combos=[2], zorro=False, eao=True,
                                        no_fusion=True, mean_pool=True)
# infer_accel_gpu.py's model (configs/CMU_config1_z_12i.yaml) has the CMU_config1_z geometry
NAMED_CONFIGS["CMU_config1_z_12i"] = dict(NAMED_CONFIGS["CMU_config1_z"])


def named_config(name: str) -> Config:
    return training_config(copy.deepcopy(NAMED_CONFIGS[name]))


def tiny_config(kind: str = "cmu", zorro: bool = False, fcl: bool = True, bimodal: bool = False,
                non_fusion_fcl: bool = False, layers: int = 2, batch_size: int = 8, eao: bool = False) -> Config:
    """Reduced token counts (same d=512 geometry) so the CPU oracle finishes in seconds."""
    if kind == "cmu":
        enc = {
            "COVAREP": {"type": "EmbeddedSequenceEncoder", "input_size": 74, "max_tokens": 150},
            "FACET": {"type": "EmbeddedSequenceEncoder", "input_size": 35, "max_tokens": 45},
            "OpenFace": {"type": "EmbeddedSequenceEncoder", "input_size": 713, "max_tokens": 70},
            "glove_vectors": {"type": "EmbeddedSequenceEncoder", "input_size": 300, "max_tokens": 20},
        }
    elif kind == "mixed":
        # the three encoder types no shipped config uses next to one EmbeddedSequenceEncoder
        enc = {
            "text": {"type": "SequenceEncoder", "num_embeddings": 60, "max_tokens": 40, "padding_idx": 0},
            "cells": {"type": "SparseTabularEncoder", "num_embeddings": 48, "max_tokens": 28, "padding_idx": 0,
                      "max_value": 100},
            "spectrogram": {"type": "PatchEncoder", "patch_size": [4, 8], "mode": "matrix", "max_tokens": 24,
                            "dropout": 0.0, "height": 16, "width": 48},
            "COVAREP": {"type": "EmbeddedSequenceEncoder", "input_size": 74, "max_tokens": 50},
        }
    else:
        enc = {
            "gene": {"type": "TabularEncoder", "num_embeddings": 90, "max_tokens": 90, "max_value": 100},
            "protein": {"type": "TabularEncoder", "num_embeddings": 37, "max_tokens": 37, "max_value": 100},
            "methylation": {"type": "TabularEncoder", "num_embeddings": 130, "max_tokens": 130, "max_value": 100},
            "mirna": {"type": "TabularEncoder", "num_embeddings": 64, "max_tokens": 64, "max_value": 100},
        }
    extra = {}
    if eao:  # the shipped *_EAO configs (configs/CMU_config1_EAO.yaml:17-28): pair passes, mean pooling, no fusion tokens
        extra = dict(eao=True, no_fusion=True, mean_pool=True, fusion_combos=[2], fcl_root=[0, 1])
    return training_config(dict(_COMMON, encoder_configs=enc, num_fusion_tokens=22, layers=layers,
                                batch_size=batch_size, bimodal_contrastive=bimodal, non_fusion_fcl=non_fusion_fcl,
                                fcl=fcl, zorro=zorro, **extra))
