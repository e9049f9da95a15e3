"""Attribute-bag configuration accepted by the drop-in models.

The reference's experiment scripts build a `Config(dataset_name)`, overwrite fields, and hand the object to
the model constructors (reference config.py:10-104, scripts/train_volleyball_stage2_dynamic.py:5-51).  The
models only *read* attributes, so this class just has to provide the same attribute names with the same
defaults; they are kept in tables below rather than as a long assignment list.  Fields the DIN stage-2 path
reads are listed in SURVEY.md §8b.
"""
import os
import time

_VOLLEYBALL_TRAIN = [1, 3, 6, 7, 10, 13, 15, 16, 18, 22, 23, 31, 32, 36, 38, 39, 40, 41, 42, 48, 50, 52, 53, 54,
                     0, 2, 8, 12, 17, 19, 24, 26, 27, 28, 30, 33, 46, 49, 51]
_VOLLEYBALL_TEST = [4, 5, 9, 11, 14, 20, 21, 25, 29, 34, 35, 37, 43, 44, 45, 47]
_COLLECTIVE_TEST = [5, 6, 7, 8, 9, 10, 11, 15, 16, 25, 28, 29]

_DEFAULTS = {
    # input / loader
    "image_size": (720, 1280), "batch_size": 32, "test_batch_size": 8, "num_boxes": 12,
    "num_frames": 3, "num_before": 5, "num_after": 4,
    # devices
    "use_gpu": True, "use_multi_gpu": True, "device_list": "0,1,2,3",
    # backbone + RoIAlign
    "backbone": "res18", "crop_size": (5, 5), "train_backbone": False, "out_size": (87, 157),
    "emb_features": 1056,
    # heads
    "num_actions": 9, "num_activities": 8, "actions_loss_weight": 1.0, "actions_weights": None,
    "num_features_boxes": 1024, "num_features_relation": 256, "num_graph": 16, "gcn_layers": 1,
    "tau_sqrt": False, "pos_threshold": 0.2,
    # optimisation
    "train_random_seed": 0, "train_learning_rate": 1e-4, "lr_plan": {11: 3e-5, 21: 1e-5},
    "train_dropout_prob": 0.3, "weight_decay": 0, "max_epoch": 30, "test_interval_epoch": 1,
    # experiment bookkeeping
    "training_stage": 1, "stage1_model_path": "", "test_before_train": False,
    "exp_note": "Group-Activity-Recognition", "exp_name": None, "set_bn_eval": False,
    "inference_module_name": "dynamic_volleyball",
    # dynamic inference (DIN)
    "stride": 1, "ST_kernel_size": 3, "dynamic_sampling": True, "sampling_ratio": [1, 3], "group": 1,
    "scale_factor": True, "beta_factor": True, "load_backbone_stage2": False, "parallel_inference": False,
    "hierarchical_inference": False, "lite_dim": None, "num_DIM": 1, "load_stage2model": False,
    "stage2model": None,
    # knobs of other heads, kept so foreign scripts can still set them
    "temporal_pooled_first": False, "halting_penalty": 0.0001,
}


class Config(object):
    def __init__(self, dataset_name):
        if dataset_name not in ("volleyball", "collective"):
            raise AssertionError(dataset_name)
        self.dataset_name = dataset_name
        for key, value in _DEFAULTS.items():
            setattr(self, key, value.copy() if isinstance(value, (list, dict)) else value)
        self.num_features_gcn = self.num_features_boxes
        if dataset_name == "volleyball":
            self.data_path = "data/volleyball/videos"
            self.train_seqs, self.test_seqs = list(_VOLLEYBALL_TRAIN), list(_VOLLEYBALL_TEST)
        else:
            self.data_path = "data/collective"
            self.test_seqs = list(_COLLECTIVE_TEST)
            self.train_seqs = [s for s in range(1, 45) if s not in self.test_seqs]

    def init_config(self, need_new_folder=True):
        """Names (and optionally creates) result/<exp_name>/ like the reference's Config.init_config."""
        if self.exp_name is None:
            stamp = time.strftime("%Y-%m-%d_%H-%M-%S", time.localtime())
            self.exp_name = "[%s_stage%d]<%s>" % (self.exp_note, self.training_stage, stamp)
        self.result_path = os.path.join("result", self.exp_name)
        self.log_path = os.path.join(self.result_path, "log.txt")
        if need_new_folder:
            os.mkdir(self.result_path)
