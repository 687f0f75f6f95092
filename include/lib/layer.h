/*
 * layer.h -- drop-in replacement for the reference's lib/layer.h (per-sample dense layer with MSE
 * back-propagation).  Field order and types of struct Layer are those of lib/layer.h:4-15 because
 * model code builds it with positional initialisers (model/my_first_model.c:29-31).
 *
 * The activation callbacks are HOST function pointers.  The library probes each callback once on
 * a small host vector; ReLU / ReLU' / constant-slope callbacks are recognised and run fused inside
 * the dense forward / backward kernels, anything else is applied on the host to the
 * host-visible result (still a CUDA GEMV underneath).
 */
#ifndef __layer_h__
#define __layer_h__

#ifdef __cplusplus
extern "C" {
#endif

struct Matrix;

struct Layer {
	int num_nodes;
	struct Matrix* nodes;
	struct Matrix* raw_nodes;
	struct Matrix* weights;
	struct Matrix* biases;
	struct Layer* previous_layer;
	void (*activation)(float*, int);
	void (*activation_ddx)(float*, int);
	char has_previous_layer;
	char has_nodes;
};

/* lib/layer.c:6-20   raw = W . prev.nodes + b ; nodes = activation(raw) */
void feed_forward(struct Layer* l);
/* lib/layer.c:22-32 */
void free_layer_data(struct Layer l);
/* lib/layer.c:34-39 */
void load_weights_from_csv(struct Layer* l, const char* filepath);
/* lib/layer.c:41-46 */
void load_biases_from_csv(struct Layer* l, const char* filepath);
/* lib/layer.c:80-107 one SGD step on the MSE loss, all layers, pre-update weights throughout */
void back_propagate_errors(struct Layer* l, float* expectations, float learn_rate);
/* lib/layer.c:48-78  (non-static in the reference, so exported here too) */
void do_back_propagate_errors(struct Layer* l, struct Layer* next_layer,
                              struct Matrix* cost_ddx_next_layer_activation, float learn_rate);

#ifdef __cplusplus
}
#endif
#endif
