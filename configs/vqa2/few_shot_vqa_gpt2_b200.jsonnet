// Few-shot VQA2 greedy answer generation with the B200-native CLIP-prefix LM (BASELINE.json configs[3]).
//
// Drop next to the reference's `configs/vqa2/clip_cap.jsonnet`; needs the import of INTEGRATION.md section 1 in
// `src/trainers/few_shot_vqa_executor.py`.  The k-shot prompt assembly (sentinel splice of vct0.py:494-533) and the
// greedy loop (clipcap.py:387-471) run inside `ClipCaptionPrefixB200.generate`; `special_token_id` is the id of the first
// added sentinel token, as for the reference's VCT0 model.
local reference = import 'clip_cap.jsonnet';

std.mergePatch(reference, {
  "experiment_name": "few_shot_vqa_gpt2_b200",
  "model_config": {
    "ModelClass": "ClipCaptionPrefixB200",
    "model_args": {
      prefix_length: 10, clip_length: 10, prefix_size: 512, mapping_type: "mlp", num_layers: 8,
      model_version: "gpt2-medium",
    },
  },
  "test": { "batch_size": 128 },
})
