# POLL_WARP / FMAD A/B on the tuned variants
B0='{}'
B1='{"POLL_WARP": true}'
B2='{"FMAD": true}'
