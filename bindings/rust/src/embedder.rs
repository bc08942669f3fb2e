//! lib/libmemex/src/llm/b200.rs -- the forward pass behind `SentenceEmbedder::runner` (llm/embedding.rs:94-135).
//! NEVER COMPILED in the build image (no rustc); `memex::B200Encoder` (memex_b200/host/embedder.cpp) is its compiled
//! twin.  The actor, the `sync_channel(100)`, `encode` / `encode_single` and `segment_text` stay as they are; the one
//! line that changes is `model.encode(&segments)` (:109): the segments are tokenised on the host with the `tokenizers`
//! crate the file already uses (:163-195) and the padded ids go to the GPU.
use std::ffi::CString;
use std::os::raw::c_void;

use tokenizers::{PaddingParams, Tokenizer, TruncationParams};

use super::embedding::{EmbeddingError, EmbeddingsModelType};
use crate::b200::ffi;

/// Architecture of an enum variant (sentence-transformers model cards) -> the two ABI structs.
pub struct Architecture {
    pub cfg: ffi::mx_model_cfg,
    pub ext: ffi::mx_model_ext,
    pub max_seq_length: usize, // sentence_bert_config.json: what rust-bert truncates to
    pub pad_id: u32,
}

pub fn architecture_of(model: EmbeddingsModelType) -> Option<Architecture> {
    let bert = |layers, hidden, ffn, normalize, max_seq_length| Architecture {
        cfg: ffi::mx_model_cfg {
            layers, hidden, heads: 12, ffn, vocab: 30522, max_pos: 512, type_vocab: 2, ln_eps: 1e-12,
            normalize, precision: 0, max_tokens: 0,
        },
        ext: ffi::mx_model_ext::default(),
        max_seq_length,
        pad_id: 0,
    };
    Some(match model {
        EmbeddingsModelType::AllMiniLmL6V2 => bert(6, 384, 1536, 1, 256),
        EmbeddingsModelType::AllMiniLmL12V2 => bert(12, 384, 1536, 1, 128),
        EmbeddingsModelType::BertBaseNliMeanTokens => bert(12, 768, 3072, 0, 128),
        EmbeddingsModelType::AllDistilrobertaV1 => {
            let mut a = bert(6, 768, 3072, 1, 512);
            a.cfg.vocab = 50265;
            a.cfg.max_pos = 514;
            a.cfg.type_vocab = 1;
            a.cfg.ln_eps = 1e-5;
            a.ext.pos_offset = 2; // RoBERTa numbers positions from padding_idx + 1
            a.pad_id = 1;
            a
        }
        EmbeddingsModelType::DistiluseBaseMultilingualCased => {
            let mut a = bert(6, 768, 3072, 0, 128);
            a.cfg.vocab = 119547;
            a.ext.no_token_type = 1;
            a.ext.dense_out = 512;
            a.ext.dense_act = ffi::MX_ACT_TANH;
            a.ext.dense_bias = 1;
            a
        }
        EmbeddingsModelType::ParaphraseAlbertSmallV2 => {
            let mut a = bert(6, 768, 3072, 0, 100);
            a.cfg.vocab = 30000;
            a.ext.ffn_act = ffi::MX_FFN_GELU_TANH;
            a.ext.embed_dim = 128;
            a.ext.share_layers = 1;
            a
        }
        EmbeddingsModelType::SentenceT5Base => {
            // T5 v1.1 base encoder + mean pool + Dense(768 -> 768, no bias) + Normalize; weights under the HF
            // T5EncoderModel names (include/memex_b200.h, mx_model_ext.family)
            let mut a = bert(12, 768, 2048, 1, 256);
            a.cfg.vocab = 32128;
            a.cfg.type_vocab = 0;
            a.cfg.ln_eps = 1e-6;
            a.ext.family = ffi::MX_FAMILY_T5;
            a.ext.d_kv = 64;
            a.ext.rel_buckets = 32;
            a.ext.rel_max_distance = 128;
            a.ext.dense_out = 768;
            a.ext.dense_bias = 0;
            a.ext.ffn_act = ffi::MX_FFN_GELU_TANH; // gated: gelu_new(wi_0 x) * wi_1 x
            a
        }
    })
}

/// Named f32 tensors under the BERT names of include/memex_b200.h (RoBERTa checkpoints use them already; DistilBERT /
/// ALBERT names are mapped as `Weights::canonicalize` does in memex_b200/host/embedder.cpp).
pub struct Weights {
    pub names: Vec<CString>,
    pub data: Vec<Vec<f32>>,
}

pub struct B200Encoder {
    handle: *mut ffi::mx_embedder,
    pub out_dim: usize,
    pub max_seq_length: usize,
    pub pad_id: u32,
}

// owned by the embedder's one OS thread (embedding.rs:84-91)
unsafe impl Send for B200Encoder {}

impl B200Encoder {
    pub fn create(arch: &Architecture, weights: &Weights, device: i32) -> Result<Self, EmbeddingError> {
        let tensors: Vec<ffi::mx_tensor> = weights
            .names
            .iter()
            .zip(weights.data.iter())
            .map(|(n, d)| ffi::mx_tensor { name: n.as_ptr(), data: d.as_ptr(), numel: d.len() as u64 })
            .collect();
        let mut handle: *mut ffi::mx_embedder = std::ptr::null_mut();
        let rc = unsafe {
            ffi::mx_embedder_create_ex(&arch.cfg, &arch.ext, tensors.as_ptr(), tensors.len() as u32, device, &mut handle)
        };
        if rc != ffi::MX_OK {
            return Err(EmbeddingError::SetupError(ffi::last_error(std::ptr::null())));
        }
        let mut dim: u32 = 0;
        unsafe { ffi::mx_embedder_out_dim(handle, &mut dim) };
        Ok(Self { handle, out_dim: dim as usize, max_seq_length: arch.max_seq_length, pad_id: arch.pad_id })
    }

    /// `model.encode(&segments)` (embedding.rs:109): [CLS] .. [SEP] (or <s> .. </s>), truncated to the model's
    /// max_seq_length, padded to the longest of the batch; one `Vec<f32>` per segment.
    pub fn encode(&self, tokenizer: &mut Tokenizer, segments: &[String]) -> Result<Vec<Vec<f32>>, EmbeddingError> {
        if segments.is_empty() {
            return Ok(Vec::new());
        }
        let _ = tokenizer.with_truncation(Some(TruncationParams { max_length: self.max_seq_length, ..Default::default() }));
        tokenizer.with_padding(None::<PaddingParams>);
        let batch = tokenizer
            .encode_batch(segments.to_vec(), true)
            .map_err(|_| EmbeddingError::EncodingFailure(segments[0].clone()))?;
        let s = batch.iter().map(|e| e.get_ids().len()).max().unwrap_or(1).max(1);
        let mut ids = vec![self.pad_id as i32; batch.len() * s];
        let mut lens = vec![0i32; batch.len()];
        for (b, e) in batch.iter().enumerate() {
            for (i, id) in e.get_ids().iter().enumerate() {
                ids[b * s + i] = *id as i32;
            }
            lens[b] = e.get_ids().len() as i32;
        }
        let mut out = vec![0f32; batch.len() * self.out_dim];
        let rc = unsafe {
            ffi::mx_embedder_encode(self.handle, ids.as_ptr(), lens.as_ptr(), batch.len() as u32, s as u32, out.as_mut_ptr())
        };
        if rc != ffi::MX_OK {
            return Err(EmbeddingError::EncodingFailure(ffi::last_error(self.handle as *const c_void)));
        }
        Ok(out.chunks(self.out_dim).map(|c| c.to_vec()).collect())
    }
}

impl Drop for B200Encoder {
    fn drop(&mut self) {
        unsafe { ffi::mx_embedder_destroy(self.handle) };
    }
}

// ---- llm/embedding.rs:94-135: what `runner` becomes ----------------------------------------------------------
//
//     fn runner(model_config: &ModelConfig, receiver: mpsc::Receiver<Message>) -> Result<(), EmbeddingError> {
//         let arch = architecture_of(model_config.model).ok_or(EmbeddingError::SetupError("Model not supported yet".into()))?;
//         let encoder = B200Encoder::create(&arch, &load_safetensors(model_dir)?, /*device*/ 0)?;   // was :99-100
//         let mut tokenizer = Tokenizer::from_file(model_dir.join("tokenizer.json"))?;
//         for (text, segment, sender) in receiver.iter() {
//             let segments = if segment { segment_text(model_config, &text)? } else { vec![text] };   // :103-107
//             let embeddings = encoder.encode(&mut tokenizer, &segments)?;                            // was :109
//             if segments.len() != embeddings.len() { ... }                                           // :110-115, unchanged
//             ...                                                                                     // :117-131, unchanged
//         }
//     }
