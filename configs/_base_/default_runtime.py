# runtime fragment shared by the configs
cudnn_deterministic = False
cudnn_benchmark = False
enable_tb = True
grad_scaler = dict(fp16=dict(initial_scale=2**16, min_scale=1, growth_interval=1000), growth_factor=2,
                   backoff_factor=0.5, max_scale=2**24, hysteresis=2)
hybrid_zero_optimizer = dict(overlap_sync_grad=True, overlap_sync_param=False, reduce_bucket_size=512 * 1024 * 1024,
                             clip_grad_norm=1.0)
loss = dict(label_smoothing=0)
adam = dict(lr=1e-4, adam_beta1=0.9, adam_beta2=0.95, adam_beta2_c=0, adam_eps=1e-8, weight_decay=0.01)
