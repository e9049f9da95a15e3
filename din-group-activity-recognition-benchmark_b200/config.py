"""Attribute-bag configuration with the reference's field names and defaults (config.py:10-116).

The stage-2 scripts mutate these fields and hand the object to the model constructors; nothing here is
computed.  Only the fields the DIN stage-2 path reads are documented in SURVEY.md §8b.
"""
import os
import time


class Config(object):
    def __init__(self, dataset_name):
        assert dataset_name in ("volleyball", "collective")
        self.dataset_name = dataset_name
        # global / input
        self.image_size = 720, 1280
        self.batch_size = 32
        self.test_batch_size = 8
        self.num_boxes = 12
        # devices
        self.use_gpu = True
        self.use_multi_gpu = True
        self.device_list = "0,1,2,3"
        # dataset splits
        if dataset_name == "volleyball":
            self.data_path = "data/volleyball/videos"
            self.train_seqs = [1, 3, 6, 7, 10, 13, 15, 16, 18, 22, 23, 31, 32, 36, 38, 39, 40, 41, 42, 48, 50,
                               52, 53, 54, 0, 2, 8, 12, 17, 19, 24, 26, 27, 28, 30, 33, 46, 49, 51]
            self.test_seqs = [4, 5, 9, 11, 14, 20, 21, 25, 29, 34, 35, 37, 43, 44, 45, 47]
        else:
            self.data_path = "data/collective"
            self.test_seqs = [5, 6, 7, 8, 9, 10, 11, 15, 16, 25, 28, 29]
            self.train_seqs = [s for s in range(1, 45) if s not in self.test_seqs]
        # backbone
        self.backbone = "res18"
        self.crop_size = 5, 5
        self.train_backbone = False
        self.out_size = 87, 157
        self.emb_features = 1056
        # labels
        self.num_actions = 9
        self.num_activities = 8
        self.actions_loss_weight = 1.0
        self.actions_weights = None
        # sampling
        self.num_frames = 3
        self.num_before = 5
        self.num_after = 4
        # relation-model sizes
        self.num_features_boxes = 1024
        self.num_features_relation = 256
        self.num_graph = 16
        self.num_features_gcn = self.num_features_boxes
        self.gcn_layers = 1
        self.tau_sqrt = False
        self.pos_threshold = 0.2
        # optimisation
        self.train_random_seed = 0
        self.train_learning_rate = 1e-4
        self.lr_plan = {11: 3e-5, 21: 1e-5}
        self.train_dropout_prob = 0.3
        self.weight_decay = 0
        self.max_epoch = 30
        self.test_interval_epoch = 1
        # experiment
        self.training_stage = 1
        self.stage1_model_path = ""
        self.test_before_train = False
        self.exp_note = "Group-Activity-Recognition"
        self.exp_name = None
        self.set_bn_eval = False
        self.inference_module_name = "dynamic_volleyball"
        # dynamic inference
        self.stride = 1
        self.ST_kernel_size = 3
        self.dynamic_sampling = True
        self.sampling_ratio = [1, 3]
        self.group = 1
        self.scale_factor = True
        self.beta_factor = True
        self.load_backbone_stage2 = False
        self.parallel_inference = False
        self.hierarchical_inference = False
        self.lite_dim = None
        self.num_DIM = 1
        self.load_stage2model = False
        self.stage2model = None
        # other heads' knobs (kept so foreign scripts can set them)
        self.temporal_pooled_first = False
        self.halting_penalty = 0.0001

    def init_config(self, need_new_folder=True):
        if self.exp_name is None:
            stamp = time.strftime("%Y-%m-%d_%H-%M-%S", time.localtime())
            self.exp_name = "[%s_stage%d]<%s>" % (self.exp_note, self.training_stage, stamp)
        self.result_path = "result/%s" % self.exp_name
        self.log_path = "result/%s/log.txt" % self.exp_name
        if need_new_folder:
            os.mkdir(self.result_path)
