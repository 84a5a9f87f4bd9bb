"""data_prepare/ of the reference: getDirectionDiffMap, SegFix_offset_helper."""
