"""The network-topology part of the two plans the reference ships (MultiTalent_plans/MultiTalent_bs4_plans_3D.pkl and
MultiTalent_resenc_bs4_plans_3D.pkl, stage 1), as plain dicts in the plans-pickle schema consumed by
`process_plans` (nnUNetTrainer.py:326-392).  Only the fields the hot path reads are kept; a real plans pickle can be
passed to the trainers instead.  `patch_size` can be overridden (BASELINE.json benchmarks 192x160x128).
"""
import copy

_GENERIC = {
    'num_stages': 2, 'num_modalities': 1, 'modalities': {0: 'CT'}, 'normalization_schemes': {0: 'CT'},
    'num_classes': 47, 'base_num_features': 30, 'conv_per_stage': 2,
    'transpose_forward': [0, 1, 2], 'transpose_backward': [0, 1, 2], 'data_identifier': 'MultiTalent_data',
    'plans_per_stage': {
        1: {'batch_size': 4, 'num_pool_per_axis': [4, 5, 5], 'patch_size': [96, 192, 192],
            'current_spacing': [1.5, 1.0, 1.0], 'original_spacing': [1.5, 1.0, 1.0], 'do_dummy_2D_data_aug': False,
            'pool_op_kernel_sizes': [[2, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2], [1, 2, 2]],
            'conv_kernel_sizes': [[3, 3, 3]] * 6},
    },
    # CT intensity statistics of the plan (dataset_properties.intensityproperties[0]) used by the synthetic generator
    'ct_clip': (-927.0, 275.0), 'ct_mean': 63.437, 'ct_sd': 175.481,
}

_RESENC = {
    'num_stages': 2, 'num_modalities': 1, 'modalities': {0: 'CT'}, 'normalization_schemes': {0: 'CT'},
    'num_classes': 47, 'base_num_features': 30, 'conv_per_stage': 2,
    'transpose_forward': [0, 1, 2], 'transpose_backward': [0, 1, 2], 'data_identifier': 'MultiTalent_data',
    'plans_per_stage': {
        1: {'batch_size': 2, 'num_pool_per_axis': [4, 5, 5], 'patch_size': [96, 192, 192],
            'current_spacing': [1.5, 1.0, 1.0], 'original_spacing': [1.5, 1.0, 1.0], 'do_dummy_2D_data_aug': False,
            'pool_op_kernel_sizes': [[1, 1, 1], [1, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2]],
            'conv_kernel_sizes': [[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]],
            'num_blocks_encoder': (1, 2, 3, 4, 4, 4), 'num_blocks_decoder': (1, 1, 1, 1, 1)},
    },
    'ct_clip': (-927.0, 275.0), 'ct_mean': 63.437, 'ct_sd': 175.481,
}


def default_plans(kind="generic", patch_size=None, batch_size=None):
    p = copy.deepcopy(_GENERIC if kind == "generic" else _RESENC)
    st = p['plans_per_stage'][1]
    if patch_size is not None:
        st['patch_size'] = list(patch_size)
    if batch_size is not None:
        st['batch_size'] = int(batch_size)
    return p
