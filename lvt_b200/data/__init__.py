from .slices import (prepare_slices, sample_abc, slice_mask, ss_shift, subscale_order,  # noqa: F401
                     visible_abc_mask, synthetic_latent_video, synthetic_vt_batch)
