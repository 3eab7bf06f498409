"""fvcore-style Checkpointer (the reference saves one per sub-network: OUTPUT_DIR/{netE,netG,netC}/
model_{iter:07d}.pth + last_checkpoint, files of the form {"model": state_dict}; meta_arch/ae.py:231-238)."""
import os

import torch


class Checkpointer:
    def __init__(self, model, save_dir="", save_to_disk=True):
        self.model, self.save_dir, self.save_to_disk = model, save_dir, save_to_disk

    def save(self, name, **extra):
        if not self.save_dir or not self.save_to_disk:
            return
        os.makedirs(self.save_dir, exist_ok=True)
        path = os.path.join(self.save_dir, f"{name}.pth")
        data = {"model": {k: v.detach().cpu() for k, v in self.model.state_dict().items()}}
        data.update(extra)
        torch.save(data, path)
        with open(os.path.join(self.save_dir, "last_checkpoint"), "w") as f:
            f.write(os.path.basename(path))

    def has_checkpoint(self):
        return os.path.exists(os.path.join(self.save_dir, "last_checkpoint"))

    def get_checkpoint_file(self):
        with open(os.path.join(self.save_dir, "last_checkpoint")) as f:
            return os.path.join(self.save_dir, f.read().strip())

    def load(self, path):
        if not path:
            return {}
        data = torch.load(path, map_location="cpu")
        state = data.pop("model") if "model" in data else data
        res = self.model.load_state_dict(state, strict=False)
        missing = list(getattr(res, "missing_keys", []) or [])
        unexpected = list(getattr(res, "unexpected_keys", []) or [])
        if missing or unexpected:  # fvcore logs both lists; a silent no-op load is the failure to avoid
            import logging
            log = logging.getLogger("lvt_b200.checkpoint")
            if missing:
                log.warning("checkpoint %s: keys missing in the file: %s", path, missing)
            if unexpected:
                log.warning("checkpoint %s: keys not used by the model: %s", path, unexpected)
            if state and len(unexpected) == len(state):
                raise KeyError(f"checkpoint {path}: none of its {len(state)} keys matches the model")
        # the bf16 / packed weight shadows of the owning engine are stale now (sub-module checkpointers wrap
        # ResEncoder / ResDecoder / DVQEmbedding, which share their parent model's engine)
        eng = getattr(self.model, "engine", None)
        if eng is not None and hasattr(eng, "shadows_fresh"):
            eng.shadows_fresh = False
        return data

    def resume_or_load(self, path, resume=True):
        if resume and self.has_checkpoint():
            path = self.get_checkpoint_file()
        return self.load(path)
