"""Modules named like the reference's, holding the drop-in callables of the geometry hot path, so that a reference
file switches over by changing its import line only:

    import postproc_other                       ->  from cdnet_b200.compat import postproc_other
    from data_prepare.getDirectionDiffMap import generate_dd_map
                                                ->  from cdnet_b200.compat.data_prepare.getDirectionDiffMap import generate_dd_map
    from data_prepare.SegFix_offset_helper import DTOffsetHelper, Sobel
                                                ->  from cdnet_b200.compat.data_prepare.SegFix_offset_helper import ...
    my_transforms_direction.LabelEncoding       ->  cdnet_b200.compat.my_transforms_direction.LabelEncoding
    my_transforms.LabelEncoding                 ->  cdnet_b200.compat.my_transforms.LabelEncoding
    from stats_utils import get_fast_aji, ...   ->  from cdnet_b200.compat.stats_utils import get_fast_aji, ...
    utils.DcmVoting2                            ->  cdnet_b200.compat.utils.DcmVoting2

Only the names on the hot path exist here; everything else stays in the reference.
"""
