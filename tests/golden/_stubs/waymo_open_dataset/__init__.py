"""Import stub (golden generator only): wosac_post_processing.py imports the Waymo protos at module level; the
tensor code we pin (`_filter_futures`, the local -> global transform of `forward`) never touches them."""
