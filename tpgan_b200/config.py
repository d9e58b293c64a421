"""Hyper-parameters of the TP-GAN hot path: the same module-level dicts, keys and values as the reference's config.py
(:29-85), restated (not imported, the reference tree is not on the GPU box).  tests/test_oracle_cpu.py asserts they are
equal to the reference's whenever /root/reference is present."""

# config.py:31-35
optimizer_param = {"learning_rate": 5e-4, "momentum": 0.9, "nesterov": True, "weight_decay": 5e-4}
# config.py:39-40
general = {"image_max_size": 1024}
# config.py:50-57
train = {"img_list": "./img.list", "learning_rate": 1e-4, "num_epochs": 50, "batch_size": 50, "log_step": 1000,
         "resume_model": None, "resume_optimizer": None}
# config.py:60-64
G = {"zdim": 64, "use_residual_block": False, "use_batchnorm": False, "num_classes": 347}
# config.py:67-68
D = {"use_batchnorm": False}
# config.py:71-82
loss = {"weight_gradient_penalty": 10, "weight_128": 1.0, "weight_64": 1.0, "weight_32": 1.5, "weight_pixelwise": 1.0,
        "weight_pixelwise_local": 3.0, "weight_symmetry": 3e-1, "weight_adv_G": 1e-3, "weight_identity_preserving": 3e1,
        "weight_total_varation": 1e-3, "weight_cross_entropy": 1e1}
# config.py:84-85
feature_extract_model = {"resume": "save/feature_extract_model/resnet18/try_1"}
