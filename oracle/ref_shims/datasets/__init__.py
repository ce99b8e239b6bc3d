"""Import shims for the reference's missing `datasets` git submodule (jramapuram/datasets@7c5d0d9).
A regular package so that it shadows both /root/reference/datasets/ (empty) and the HuggingFace `datasets`."""
