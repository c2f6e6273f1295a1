"""Import shim (test infrastructure, used only by oracle/ref_harness.py in the build container).

Restates the three timm classes the reference imports at
train_settings/dvd/improved_diffusion/cross_model.py:7 so that the UNMODIFIED reference can be
imported without the (absent, unpinned: requirements.txt:17) timm package."""
