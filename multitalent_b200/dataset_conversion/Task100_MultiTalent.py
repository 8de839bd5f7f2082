"""MultiTalent dataset / label / region tables -- the constants of
nnunet/dataset_conversion/Task100_MultiTalent.py:35-207, exposed under the reference's names.

They are DATA consumed by the loss and the predictor: 13 source datasets, 47 global labels, 47 output channels
("regions", three of which are unions of two labels, so channels overlap -> sigmoid heads).  The tables are expanded
from one compact per-dataset spec; tests/test_tables.py checks the expansion against a fixture dumped from the
reference module.
"""

# (task id, "prefix", [structure names in local-label order], {extra union regions: local labels}, region order)
# Global label ids are assigned consecutively in this order, starting at 1.
_SPEC = (
    ("Task003_Liver", "03", ("liver_wo_cancer", "liver_tumor"),
     (("liver", (1, 2)), ("cancer", (2,)))),
    ("Task006_Lung", "06", ("lung_nodule",), (("lungnodule", (1,)),)),
    ("Task007_Pancreas", "07", ("pancreas_wo_cancer", "pancreas_cancer"),
     (("pancreas", (1, 2)), ("pancreas_cancer", (2,)))),
    ("Task008_HepaticVessel", "08", ("hepatic_vessel", "liver_cancer"), (("vessel", (1,)), ("tumor", (2,)))),
    ("Task009_Spleen", "09", ("spleen",), None),
    ("Task010_Colon", "10", ("colon_cancer",), None),
    ("Task017_AbdominalOrganSegmentation", "17",
     ("spleen", "right_kidney", "left_kidney", "gallbladder", "esophagus", "liver_whole", "stomach", "aorta",
      "inf_vena_cava", "port_and_splen_vein", "pancreas_whole", "right_adrenal_gland", "left_adrenal_gland"),
     {"liver_whole": "liver", "pancreas_whole": "pancreas"}),
    ("Task046_AbdOrgSegm2", "46",
     ("spleen", "left_kidney", "gallbladder", "esophagus", "liver", "stomach", "pancreas", "duodenum"), None),
    ("Task051_StructSeg2019_Task3_Thoracic_OAR", "51",
     ("left_lung", "right_lung", "heart", "esophagus", "bronchies", "spinal_cord_nerve_thingy"), None),
    ("Task055_SegTHOR", "55", ("esophagus", "heart", "trachea", "aorta"), None),
    ("Task062_NIHPancreas", "62", ("pancreas",), None),
    ("Task064_KiTS_labelsFixed", "64", ("both_kidneys_wo_tumor", "kidney_tumor"),
     (("both_kidneys", (1, 2)), ("kidney_tumor", (2,)))),
    ("Task018_PelvicOrganSegmentation", "18", ("bladder", "uterus", "rectum", "small_bowel"), None),
)

MultiTalent_task_ids = [s[0] for s in _SPEC]
MultiTalent_task_label_maps = {}
MultiTalent_labels = {}
MultiTalent_regions = {}
MultiTalent_regions_class_order = {}
MultiTalent_valid_regions = {}

_next = 1
for _task, _pre, _names, _regions in _SPEC:
    _glob = tuple(range(_next, _next + len(_names)))
    _next += len(_names)
    MultiTalent_task_label_maps[_task] = (tuple(range(1, len(_names) + 1)), _glob)
    for _g, _n in zip(_glob, _names):
        MultiTalent_labels[_g] = "%s_%s" % (_pre, _n)
    if isinstance(_regions, tuple):     # explicit (union) regions given in local labels
        _rl = [("%s_%s" % (_pre, n), tuple(_glob[i - 1] for i in loc)) for n, loc in _regions]
    else:                               # one region per structure, optionally renamed
        _ren = _regions or {}
        _rl = [("%s_%s" % (_pre, _ren.get(n, n)), (g,)) for g, n in zip(_glob, _names)]
    for _n, _l in _rl:
        MultiTalent_regions[_n] = _l
    MultiTalent_regions_class_order[_task] = _glob
    MultiTalent_valid_regions[_task] = tuple(n for n, _ in _rl)

MultiTalent_region_output_idx_mapping = {j: i for i, j in enumerate(MultiTalent_regions.keys())}

NUM_OUTPUT_CHANNELS = len(MultiTalent_regions)
NUM_LABELS = _next  # 0 .. 47


def region_bitmasks():
    """(pos_mask, region->channel): pos_mask[label] has bit j set iff `label` belongs to the region of channel j --
    the on-device form of the `target == l` OR-loop at MultiTalent_Trainer_DDP.py:580-584."""
    pos = [0] * NUM_LABELS
    for name, labs in MultiTalent_regions.items():
        j = MultiTalent_region_output_idx_mapping[name]
        for l in labs:
            pos[l] |= 1 << j
    return pos, dict(MultiTalent_region_output_idx_mapping)


def valid_channel_mask(valid_regions):
    """Bitmask of output channels supervised for a sample whose dataset labels `valid_regions` (region names)."""
    m = 0
    for r in valid_regions:
        m |= 1 << MultiTalent_region_output_idx_mapping[r]
    return m
