F0='{}'
F1='{"FAST_PATH": false}'
F2='{"FAST_PATH": false, "UNIFORM_WARP": true}'
F3='{"UNIFORM_WARP": true}'
