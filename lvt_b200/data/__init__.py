from .slices import (prepare_slices, prepare_slices_batched, sample_abc, slice_mask, ss_shift, subscale_order,  # noqa: F401
                     visible_abc_mask, synthetic_latent_video, synthetic_vt_batch)
from .latents import (CodesExtractor, extract_codes, get_latent_video_paths, latent_slice_loader,  # noqa: F401
                      load_latent_video, save_latent_video)
