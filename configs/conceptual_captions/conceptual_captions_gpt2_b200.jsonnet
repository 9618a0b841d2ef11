// Conceptual Captions mapper training with the B200-native CLIP-prefix LM step (BASELINE.json configs[1]).
//
// Drop this file next to the reference's `configs/conceptual_captions/conceptual_captions.jsonnet` and run the
// reference's own entry point:   python main.py ../configs/conceptual_captions/conceptual_captions_gpt2_b200.jsonnet --mode train
// after adding the one import of INTEGRATION.md section 1 to `src/trainers/clipcap_exector.py`
// (`from eavqa_b200 import ClipCaptionPrefixB200`: the executor resolves `ModelClass` by name in its module globals,
// clipcap_exector.py:52-53).
//
// Everything not listed here -- data loader, ModuleParser input modules, cache, metrics, validation / test settings --
// is inherited unchanged from the reference's config.  What changes relative to it: the GPT-2 flavour of the model
// block (the reference's Conceptual Captions config ships with the T0 model; its GPT-2 keys are those of
// configs/vqa2/clip_cap.jsonnet:25-43) and the class name.
local reference = import 'conceptual_captions.jsonnet';

std.mergePatch(reference, {
  "experiment_name": "conceptual_captions_gpt2_b200",
  "model_config": {
    "base_model": "gpt2",
    "ModelClass": "ClipCaptionPrefixB200",       // was "ClipCaptionPrefix": same constructor kwargs, same surface
    "TokenizerClass": "GPT2Tokenizer",
    "TokenizerModelVersion": "gpt2",
    "ConfigClass": "GPT2Config",
    "model_args": {
      prefix_length: 10,
      clip_length: 10,
      prefix_size: 512,                            // CLIP ViT-B/32 embeddings
      mapping_type: "transformer",                 // 8-layer transformer mapper (clipcap.py:213-237); "mlp" for configs[0]
      num_layers: 8,
      model_version: "gpt2",                       // frozen LM, packed once into the engine
    },
    "SPECIAL_TOKENS": { "bos_token": "<BOS>", "additional_special_tokens": [] },
  },
  "train": {
    "type": "ClipCapExecutor",
    "batch_size": 256,                             // per GPU; Lightning DDP shards the sampler, one process per B200
    "additional": { "gradient_accumulation_steps": 1 },
  },
})
