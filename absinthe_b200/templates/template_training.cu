{#- Training programs (ref template_training.cpp) use the kernels of template_tiling.cu unchanged: the
    launcher selects the training protocol -- tiles 0, 20, 40, ... (one per SM), no periodic update --
    from the "# training 1" line of the descriptor (absinthe_variant_run, training = 1). -#}
// training variant: every 20th tile, one CTA each, no halo update (ref template_training.cpp:276-311)
{% include "template_tiling.cu" %}
