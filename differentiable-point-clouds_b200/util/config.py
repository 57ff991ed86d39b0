"""The `cfg` argument of the hot path.

The reference passes an `EasyDict` built from `dpc/resources/default_config.yaml` plus an
experiment YAML plus `--key=value` overrides (dpc/util/config.py:7-29,105-149).  easydict
is not installed, so this module provides the same attribute-style mapping with the
reference's defaults for every key the projection path reads and the
reference's merge rules: unknown keys raise KeyError (config.py:16-17), values are coerced
to the type of the default, and a bool can only be overridden by a bool (config.py:19-29).
YAML is read with `yaml.safe_load` (the reference's bare `yaml.load` fails on PyYAML >= 6).
"""
import copy

_DEFAULTS = {
    # networks (default_config.yaml:10-26,37-47)
    "image_size": 128,
    "z_dim": 1024,
    "f_dim": 16,
    "fc_dim": 1024,
    "pc_decoder_init_stddev": 0.025,
    "pc_unit_cube": True,
    "pose_candidates_num_layers": 3,
    "pose_predictor_student": True,
    "pose_predictor_student_loss_weight": 1.0,
    "pose_student_align_loss": False,
    "predict_translation_scaling_factor": 0.15,
    "predict_translation_tanh": True,
    "predict_translation_init_stddev": 0.05,
    # loss / optimisation (default_config.yaml:92-122)
    "proj_weight": 1.0,
    "drc_weight": 0.0,
    "proj_rgb_weight": 0.0,
    "proj_depth_weight": 0.0,
    "max_dataset_depth": 10.0,
    "pc_gauss_filter_gt": False,
    "pc_gauss_filter_gt_rgb": False,
    "pc_gauss_filter_gt_switch_off": False,
    "bicubic_gt_downsampling": False,
    "clip_gradient_norm": 0.0,
    "learning_rate": 0.0001,
    "learning_rate_step": 1.0,
    "learning_rate_2": 0.00001,
    "weight_decay": 0.001,
    "variable_num_views": False,
    # shapes set by the caller (default_config.yaml:25-31,107-108,37)
    "pc_num_points": 8000,
    "pc_point_dropout": 1.0,
    "pc_point_dropout_scheduled": True,
    "pc_point_dropout_exponential_schedule": False,
    "pc_point_dropout_end_step": 1.0,
    "pc_point_dropout_start_step": 0.0,
    "batch_size": 8,
    "step_size": 4,
    "pose_predict_num_candidates": 1,
    "max_number_of_steps": 600000,
    # pose (default_config.yaml:33-47)
    "predict_pose": False,
    "pose_quaternion": True,
    "predict_translation": False,
    # points -> grid (default_config.yaml:49-58)
    "pc_relative_sigma": 1.0,
    "pc_relative_sigma_end": 0.2,
    "pc_fast": True,
    "pc_gauss_kernel_size": 11,
    "pc_separable_gauss_filter": True,
    "pc_learn_occupancy_scaling": True,
    "pc_occupancy_scaling_maximum": 1.0,
    # rgb (default_config.yaml:60-66)
    "pc_rgb": False,
    "pc_rgb_stop_points_gradient": False,
    "pc_rgb_clip_after_conv": False,
    "pc_rgb_divide_by_occupancies": False,
    "pc_rgb_divide_by_occupancies_epsilon": 0.01,
    "pc_rgb_deep_decoder": False,
    "learn_focal_length": False,
    "focal_length_range": 1.0,
    "focal_length_mean": 2.0,
    # projection (default_config.yaml:77-90)
    "vox_size": 64,
    "vox_size_z": -1,
    "focal_length": 1.875,
    "camera_distance": 2.0,
    "ptn_max_projection": False,
    "max_depth": 10.0,
    "drc_logsum": True,
    "drc_logsum_clip_val": 0.00001,
    "drc_tf_cumulative": True,
}


class AttrDict(dict):
    """dict with attribute access (stand-in for easydict.EasyDict)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def copy(self):
        return AttrDict(copy.deepcopy(dict(self)))


def _coerce(key, old, new):
    if isinstance(old, bool):
        if not isinstance(new, bool):
            raise ValueError("config key %r is a bool and can only be overridden by a bool, got %r" % (key, new))
        return new
    if isinstance(old, float) and isinstance(new, (int, float)) and not isinstance(new, bool):
        return float(new)
    if isinstance(old, int) and isinstance(new, int) and not isinstance(new, bool):
        return int(new)
    if type(old) is type(new):
        return new
    raise ValueError("type mismatch for config key %r: default %r (%s) vs %r (%s)"
                     % (key, old, type(old).__name__, new, type(new).__name__))


def default_config(**overrides):
    """Reference defaults for the projection path, then `overrides` merged with the
    reference's rules (KeyError for an unknown key)."""
    cfg = AttrDict(copy.deepcopy(_DEFAULTS))
    merge_into(cfg, overrides)
    return cfg


def merge_into(cfg, overrides):
    for k, v in overrides.items():
        if k not in cfg:
            raise KeyError("%s is not a valid config key" % k)
        cfg[k] = _coerce(k, cfg[k], v)
    return cfg


def experiment_config(name):
    """The two experiment overrides that exist in the reference
    (experiments/chair_camera_supervision/config.yaml, experiments/chair_unsupervised/config.yaml)."""
    common = dict(vox_size=64, pc_gauss_kernel_size=21, pc_relative_sigma=3.0,
                  pc_num_points=8000, pc_point_dropout=0.07)
    if name == "chair_camera_supervision":
        return default_config(**common)
    if name == "chair_unsupervised":
        return default_config(predict_pose=True, pose_predict_num_candidates=4,
                              pose_predictor_student_loss_weight=20.0, **common)
    raise KeyError(name)


# Keys of the reference's default_config.yaml that configure subsystems outside this package (data input, checkpoints,
# visualisation, evaluation drivers, Blender rendering): accepted in a YAML file and dropped.
_IGNORED_KEYS = frozenset("""
config inp_dir synth_set num_views num_views_to_use tfrecords_gzip_compressed saved_camera saved_depth
encoder_name decoder_name posenet_name decoder_conv_init_stdev focal_range voxel_grid_size checkpoint_dir gpu
gpu_allow_growth per_process_gpu_memory_fraction shuffle_batch shuffle_dataset drc_rgb_weight compute_validation_loss
validation_interval save_intermediate_pcs save_intermediate_pcs_interval num_dataset_samples save_predictions_dir
vis_threshold vox_threshold vis_size vis_voxels vis_depth_projs vis_all_views save_individual_images save_predictions
save_voxels save_point_clouds save_as_mat save_rotated_points models_list gt_pc_dir eval_split pc_eval_chamfer_num_parts
eval_unsupervised_shape save_val_projection save_val_projection_dir vox_marching_cubes_isosurface
vox_marching_cubes_dense pose_accuracy_threshold vis_azimuth vis_elevation vis_dist render_image_size
render_cycles_samples render_colored_subsets
""".split())
# Keys that change what is computed and have NO implementation here: a YAML may carry them only at the reference default.
_UNIMPLEMENTED_AT_DEFAULT = {"pc_normalise_gauss": False, "pc_normalise_gauss_analytical": True, "align_to_canonical": False,
                             "pc_unit_cube": True, "pc_fast": True, "pc_rgb_deep_decoder": False}


def load_config(path, **overrides):
    """Defaults <- YAML file <- overrides, with the reference's rule that an unknown key is an error
    (config.py:16-17).  Keys of subsystems outside this package (_IGNORED_KEYS) are dropped; a key that changes the
    objective but is not implemented here raises unless it has the reference's default value."""
    import yaml
    with open(path) as f:
        data = yaml.safe_load(f) or {}
    cfg = default_config()
    take = {}
    for k, v in data.items():
        if k in _UNIMPLEMENTED_AT_DEFAULT and v != _UNIMPLEMENTED_AT_DEFAULT[k]:
            raise NotImplementedError("config key %s=%r is not implemented in dpc_b200 (only the reference default %r is)"
                                      % (k, v, _UNIMPLEMENTED_AT_DEFAULT[k]))
        if k in cfg:
            take[k] = v
        elif k not in _IGNORED_KEYS and k not in _UNIMPLEMENTED_AT_DEFAULT:
            raise KeyError("%s is not a valid config key" % k)
    merge_into(cfg, take)
    merge_into(cfg, overrides)
    return cfg
