"""Default settings, same structure and key names as benchmark/wifi_csi/preset.py:8-96.

Only keys the THAT path consumes are meaningful here; keys of other reference models (DETR loss, scheduler,
object queries) are kept so that code indexing them keeps working.  New OPTIONAL keys -- ``nn.dtype`` ("bf16" |
"fp32") and ``nn.world_size`` -- default to values that reproduce the reference call sequence.
"""
preset = {
    "model": "THAT",
    "task": "activity",                                  # "identity", "activity", "location"
    "repeat": 8,
    "path": {
        "data_x": "dataset/wifi_csi/amp",                 # directory of CSI amplitude files (*.npy)
        "data_y": "dataset/annotation.csv",               # annotation file
        "save": "results/result.json",
    },
    "data": {
        "num_users": ["0", "1", "2", "3", "4", "5"],
        "wifi_band": ["5"],
        "environment": ["empty_room"],
        "length": 3000,
    },
    "data_band2": {
        "num_users": ["0", "1", "2", "3", "4", "5"],
        "wifi_band": ["5"],
        "environment": ["empty_room"],
        "length": 3000,
    },
    "nn": {
        "lr": 5e-4,
        "epoch": 300,
        "batch_size": 16,
        "threshold": 0.5,
        "scheduler": {"type": "cosine_warmup", "num_warmup_epochs": 10, "min_lr_ratio": 0.05},
        "loss": {"type": "HungarianMatchingLoss", "cost_class_weight": 1.0, "aux_loss_weight": 0.25,
                 "label_smoothing": 0.3, "class_imbalance_weight": 0.25},
        "cross_attention_temp": 2,
        "weight_decay": 2e-4,
        "num_obj_queries": 5,
        "num_decoder_layers": 6,
        "dim_FFN": 512,
        "token_length": 10,
        # --- new, optional (defaults reproduce the reference behaviour) ---
        "dtype": "bf16",                                 # compute type of the contractions: "bf16" | "fp32"
        "world_size": 1,                                 # data-parallel ranks (one process per GPU)
    },
    "encoding": {
        "activity": {
            "nan":      [0, 0, 0, 0, 0, 0, 0, 0, 0],
            "nothing":  [1, 0, 0, 0, 0, 0, 0, 0, 0],
            "walk":     [0, 1, 0, 0, 0, 0, 0, 0, 0],
            "rotation": [0, 0, 1, 0, 0, 0, 0, 0, 0],
            "jump":     [0, 0, 0, 1, 0, 0, 0, 0, 0],
            "wave":     [0, 0, 0, 0, 1, 0, 0, 0, 0],
            "lie_down": [0, 0, 0, 0, 0, 1, 0, 0, 0],
            "pick_up":  [0, 0, 0, 0, 0, 0, 1, 0, 0],
            "sit_down": [0, 0, 0, 0, 0, 0, 0, 1, 0],
            "stand_up": [0, 0, 0, 0, 0, 0, 0, 0, 1],
        },
        "location": {
            "nan": [0, 0, 0, 0, 0],
            "a":   [1, 0, 0, 0, 0],
            "b":   [0, 1, 0, 0, 0],
            "c":   [0, 0, 1, 0, 0],
            "d":   [0, 0, 0, 1, 0],
            "e":   [0, 0, 0, 0, 1],
        },
    },
    "pretrained_path": None,
    "transfer_scenario": "full",
    "save_model": False,
    "saving_path": "results/",
}
