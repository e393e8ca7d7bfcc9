COL='"_clean":true,"_assignment":[0,0,0,0,0,0,0,0,0],"STREAM_K":"regs","SMEM_BUDGET":232448,"FUSE_HALO":true,"WAIT_SLEEP":100,"_shapes":[[57,9,80]],"BLOCK":[1,1]'
c() { echo "{$COL,\"COLUMN\":{\"WARPS\":${2:-[2,9]},\"UNROLL\":1,\"SLOTS\":3,\"MERGE\":true$1}}"; }
CANDS=('{}'
 "$(c '')"
 "$(c ',"CLUSTER":[2,1]')"
 "$(c ',"CLUSTER":[4,1]')"
 "$(c ',"CLUSTER":[8,1]')"
 "$(c ',"CLUSTER":[2,2]')"
 "$(c ',"CLUSTER":[4,2]')"
 "$(c ',"CLUSTER":[2,1],"SLACK":2')"
 "$(c ',"CLUSTER":[2,1],"SLACK":8')"
 "$(c ',"CLUSTER":[2,1]' '[1,18]')"
 "$(c ',"CLUSTER":[4,1]' '[1,18]')"
 "$(c ',"CLUSTER":[8,1]' '[1,18]')"
 "$(c ',"CLUSTER":[2,1]' '[2,10]')")
