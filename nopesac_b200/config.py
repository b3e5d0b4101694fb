"""Config surface of the path: `get_cfg()` (the detectron2 0.4 defaults the reference YAMLs touch) and
`get_sparseplane_cfg_defaults(cfg)` (same keys and values as NopeSAC_Net/config/config.py:5-114), so that
`configs/inference_mp3d.yaml` + `Base.yaml` of the reference load unchanged:

    cfg = get_cfg(); get_sparseplane_cfg_defaults(cfg)
    cfg.merge_from_file(".../configs/inference_mp3d.yaml"); cfg.merge_from_list(opts); cfg.freeze()

(test_NopeSAC.py:182-188).  Keys actually read by the hot path are listed in SURVEY.md §5.
"""
from __future__ import annotations

from .compat import CfgNode

# detectron2 defaults (detectron2/config/defaults.py @0.4) for the sub-tree the reference configs use
_D2_DEFAULTS = {
    "VERSION": 2,
    "MODEL": {
        "LOAD_PROPOSALS": False, "MASK_ON": False, "KEYPOINT_ON": False, "DEVICE": "cuda",
        "META_ARCHITECTURE": "GeneralizedRCNN", "WEIGHTS": "",
        "PIXEL_MEAN": [103.530, 116.280, 123.675], "PIXEL_STD": [1.0, 1.0, 1.0],
        "BACKBONE": {"NAME": "build_resnet_backbone", "FREEZE_AT": 2},
        "RESNETS": {
            "DEPTH": 50, "OUT_FEATURES": ["res4"], "NUM_GROUPS": 1, "NORM": "FrozenBN", "WIDTH_PER_GROUP": 64,
            "STRIDE_IN_1X1": True, "RES5_DILATION": 1, "RES2_OUT_CHANNELS": 256, "STEM_OUT_CHANNELS": 64,
            "DEFORM_ON_PER_STAGE": [False, False, False, False], "DEFORM_MODULATED": False, "DEFORM_NUM_GROUPS": 1,
        },
        "SEM_SEG_HEAD": {
            "NAME": "SemSegFPNHead", "IN_FEATURES": ["p2", "p3", "p4", "p5"], "IGNORE_VALUE": 255,
            "NUM_CLASSES": 54, "CONVS_DIM": 128, "COMMON_STRIDE": 4, "NORM": "GN", "LOSS_WEIGHT": 1.0,
        },
    },
    "INPUT": {"FORMAT": "BGR", "MIN_SIZE_TEST": 800, "MAX_SIZE_TEST": 1333, "MASK_FORMAT": "polygon"},
    "DATASETS": {"TRAIN": (), "TEST": ()},
    "DATALOADER": {"NUM_WORKERS": 4, "ASPECT_RATIO_GROUPING": True, "FILTER_EMPTY_ANNOTATIONS": True,
                   "SAMPLER_TRAIN": "TrainingSampler"},
    "SOLVER": {
        "LR_SCHEDULER_NAME": "WarmupMultiStepLR", "MAX_ITER": 40000, "BASE_LR": 0.001, "MOMENTUM": 0.9,
        "NESTEROV": False, "WEIGHT_DECAY": 0.0001, "WEIGHT_DECAY_NORM": 0.0, "GAMMA": 0.1, "STEPS": (30000,),
        "WARMUP_FACTOR": 1.0 / 1000, "WARMUP_ITERS": 1000, "WARMUP_METHOD": "linear", "CHECKPOINT_PERIOD": 5000,
        "IMS_PER_BATCH": 16, "BIAS_LR_FACTOR": 1.0, "WEIGHT_DECAY_BIAS": 0.0001,
        "CLIP_GRADIENTS": {"ENABLED": False, "CLIP_TYPE": "value", "CLIP_VALUE": 1.0, "NORM_TYPE": 2.0},
    },
    "TEST": {"EXPECTED_RESULTS": [], "EVAL_PERIOD": 0, "DETECTIONS_PER_IMAGE": 100},
    "OUTPUT_DIR": "./output",
    "SEED": -1,
    "CUDNN_BENCHMARK": False,
    "VIS_PERIOD": 0,
}


def get_cfg() -> CfgNode:
    return CfgNode(_D2_DEFAULTS).clone()


_NOPESAC_DEFAULTS = {
    "SOLVER": {"WEIGHT_DECAY_EMBED": 0.0, "OPTIMIZER": "ADAMW", "BACKBONE_MULTIPLIER": 1.0,
               "SEM_SEG_HEAD_MULTIPLIER": 1.0, "PLANE_MATCHER_HEAD_MULTIPLIER": 1.0},
    "MODEL": {
        "FREEZE": [], "DEPTH_ON": False, "EMBEDDING_ON": False, "CAMERA_ON": False, "MASK_ON": True,
        "HUNGARIAN_MATCHER_ON": True, "LOSS_DETECTION_ON": True, "LOSS_CAMERA_ON": False, "LOSS_EMB_ON": False,
        "SEM_SEG_HEAD": {
            "DEEP_SUPERVISION": True, "NO_OBJECT_WEIGHT": 0.1, "DICE_WEIGHT": 1.0, "MASK_WEIGHT": 20.0,
            "PARAM_WEIGHT_L1": 0.5, "PARAM_WEIGHT_COS": 10.0, "PARAM_HM_WEIGHT_L1": 0.5, "PARAM_WEIGHT_Q": 1.0,
            "PARAM_WEIGHT_CENTER_INS": 0.5, "PARAM_WEIGHT_ANGLE": 0.0028, "PARAM_WEIGHT_OFFSET": 0.01,
            "NUM_CLASSES": 1, "CENTER_ON": False, "PARAM_ON": False, "PARAM_IN_MATCHER": True, "NHEADS": 8,
            "ENC_LAYERS": 6, "DEC_LAYERS": 6, "NUM_OBJECT_QUERIES": 50, "MASK_DIM": 256, "HIDDEN_DIM": 256,
        },
        "CAMERA_BRANCH": "CACHED",
        "CAMERA_HEAD": {
            "NAME": "", "LOSS_WEIGHT": 1.0, "KMEANS_TRANS_PATH": "./camCls/kmeans_trans_32.pkl",
            "KMEANS_ROTS_PATH": "./camCls/kmeans_rots_32.pkl", "TRANS_CLASS_NUM": 32, "ROTS_CLASS_NUM": 32,
            "FEATURE_SIZE": 64, "BACKBONE_FEATURE": "res3", "REFINE_ON": False, "CAM_REC_ON": False,
            "RAND_ON": False, "PIXEL_CAM_FIX_ON": False, "INFERENCE_OUT_CAM_TYPE": "soft",
            "INITIAL_CAM_WEIGHT": 1.0, "PLANE_CAM_WEIGHT": 1.0, "PLANE_CAM_WEIGHT_PREDPLANE": 0.1,
            "CLASSIFICATION_ON": False, "INFERENCE_SP_TOPCAM_ON": False, "INFERENCE_SP_TOPCAM_PATH": "",
            "WARP_PLANE_IN_CAM_REF_ON": True,
        },
        "MATCHING_HEAD": {"NAME": "", "INITIAL_CAM_ON": True, "OFFSET_MULTIPLIER": 4.0, "NORMAL_MULTIPLIER": 8.0},
    },
    "TEST": {
        "EVAL_GT_BOX": False, "OVERLAP_THRESHOLD": 0.6, "PLANE_SCORE_THRESHOLD": 0.6, "MASK_PROB_THRESHOLD": 0.5,
        "EVAL_FULL_SCENE": False, "MATCHING_SCORE_THRESHOLD": 0.2, "POSE_REFINEMENT_WITH_GT_MATCHERS": False,
        "POSE_REFINEMENT_WITH_GT_NOISE_MATCHERS": False,
        "POSE_REFINEMENT_WITH_GT_NOISE_MATCHERS_OFFSET_SCALE": 0.1,
        "POSE_REFINEMENT_WITH_GT_NOISE_MATCHERS_NORMAL_SCALE": 10.0,
    },
    "DATALOADER": {"ASPECT_RATIO_GROUPING": False, "AUGMENTATION": False},
    "DEBUG_ON": False, "DEBUG_CAMERA_ON": False, "SEED": 42, "FIX_SEED": True,
    "DATASETS": {"ROOT_DIR": ""},
}


def get_sparseplane_cfg_defaults(cfg: CfgNode) -> CfgNode:
    """Adds the NopeSAC keys onto a detectron2-style cfg (same name as the reference function)."""
    cfg.merge_from_other_cfg(CfgNode(_NOPESAC_DEFAULTS))
    return cfg


def load_config(config_file: str, opts=()) -> CfgNode:
    """test_NopeSAC.py:182-188 `setup()` without the logger."""
    cfg = get_cfg()
    get_sparseplane_cfg_defaults(cfg)
    cfg.merge_from_file(config_file)
    cfg.merge_from_list(list(opts))
    cfg.freeze()
    return cfg


def inference_cfg(num_queries: int = 50, out_cam_type: str = "soft", match_threshold: float = 0.2,
                  device: str = "cuda") -> CfgNode:
    """The values `configs/inference_mp3d.yaml` (+ Base.yaml) sets, without needing the reference tree on
    disk (benches / GPU tests); NUM_OBJECT_QUERIES etc. overridable like `opts` on the reference CLI."""
    cfg = get_cfg()
    get_sparseplane_cfg_defaults(cfg)
    cfg.merge_from_other_cfg(CfgNode({
        "MODEL": {
            "META_ARCHITECTURE": "PlaneTR_NopeSAC", "MASK_ON": True, "CAMERA_ON": True, "EMBEDDING_ON": True,
            "DEVICE": device, "PIXEL_MEAN": [123.675, 116.280, 103.530], "PIXEL_STD": [58.395, 57.120, 57.375],
            "BACKBONE": {"FREEZE_AT": 0},
            "RESNETS": {"STRIDE_IN_1X1": False, "OUT_FEATURES": ["res2", "res3", "res4", "res5"]},
            "SEM_SEG_HEAD": {"NAME": "PlaneTRHead", "IN_FEATURES": ["res2", "res3", "res4", "res5"], "NORM": "GN",
                             "NUM_CLASSES": 1, "PARAM_ON": True, "CENTER_ON": True,
                             "NUM_OBJECT_QUERIES": num_queries},
            "CAMERA_HEAD": {"REFINE_ON": True, "CAM_REC_ON": True, "INFERENCE_OUT_CAM_TYPE": out_cam_type,
                            "NAME": "PlaneCameraHead", "WARP_PLANE_IN_CAM_REF_ON": True},
        },
        "INPUT": {"FORMAT": "RGB"},
        "TEST": {"MATCHING_SCORE_THRESHOLD": match_threshold},
    }))
    cfg.freeze()
    return cfg
