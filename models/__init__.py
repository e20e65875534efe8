"""Import-path shim: the reference pickles whole models (train.py:222), so its class paths
``models.bidate_model.BiDateNet`` and ``models.unet_parts.*`` must resolve.  They resolve to fabric_b200."""
