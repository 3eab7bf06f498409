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
        self.model.load_state_dict(state, strict=False)
        return data

    def resume_or_load(self, path, resume=True):
        if resume and self.has_checkpoint():
            path = self.get_checkpoint_file()
        return self.load(path)
