#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *real* reference (damians13/big-linear-algebra) as the
# parity oracle.  Nothing under oracle/ is ever linked, imported or executed by the product
# (libbla.so); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
# legs may use it.
#
# The reference sources are compiled from where they lie (REF=/root/reference, read-only); no
# reference source is copied into the repository.  Outputs go ONLY to oracle/_ref/ (git-ignored,
# but it travels to the GPU box with the gpurun snapshot).  Variants:
#   libref_f64.so          HEAD as shipped (matrix_float_t = double, lib/matrix.h:4)
#   libref_f32.so          typedef patched to float (needed by the float-era layer.c, D1)
#   libref_{f64,f32}_convfix.so   + the two swapped assignment lines of lib/conv.c:183/199
#                          exchanged (D3) -- without it conv() never writes its output
# The patches are the sed one-liners below, applied to a symlink tree under oracle/_ref/gen that
# is removed after the build.
set -euo pipefail
REF="${REF:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/lib" ]; then
  echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt $OUT" >&2
  exit 0
fi
CC="${CC:-gcc}"
CFLAGS="-O2 -std=c99 -fPIC -w -fno-fast-math -ffp-contract=off"
mkdir -p "$OUT"
GEN="$OUT/gen"
rm -rf "$GEN"

mk_tree() {            # $1 = variant name, $2 = float|double, $3 = convfix? (0/1)
  local t="$GEN/$1"
  mkdir -p "$t/lib" "$t/model"
  for f in "$REF"/lib/*; do ln -s "$f" "$t/lib/$(basename "$f")"; done
  for f in "$REF"/model/*.c "$REF"/main.c; do ln -s "$f" "$t/model/$(basename "$f")"; done
  if [ "$2" = float ]; then
    rm "$t/lib/matrix.h"
    sed 's/^typedef double matrix_float_t;/typedef float matrix_float_t;/' "$REF/lib/matrix.h" > "$t/lib/matrix.h"
    grep -q 'typedef float matrix_float_t' "$t/lib/matrix.h"
  fi
  if [ "$3" = 1 ]; then
    rm "$t/lib/conv.c"
    # D3: exchange the bodies' assignment lines so each reshape does what its name says
    sed -e '183s/.*/                matrix->data[index * num_channels + c] = channels[c].data[index];/' \
        -e '199s/.*/                channels[c].data[index] = matrix->data[index * num_channels + c];/' \
        "$REF/lib/conv.c" > "$t/lib/conv.c"
  fi
}

build_lib() {          # $1 = variant, $2 = with layer.c? (0/1)
  local t="$GEN/$1"
  local srcs="$t/lib/matrix.c $t/lib/conv.c $t/lib/norm.c $t/lib/util.c $t/lib/csv.c"
  if [ "$2" = 1 ]; then srcs="$srcs $t/lib/layer.c"; fi
  $CC $CFLAGS -shared -Wl,-Bsymbolic -o "$OUT/libref_$1.so" $srcs -lm
}

mk_tree f64 double 0;          build_lib f64 0
# the reference's data readers / samplers (lib/mnist_csv2.c, lib/cifar10.c) for the data-pipeline parity tests
$CC $CFLAGS -shared -Wl,-Bsymbolic -o "$OUT/libref_data.so" "$GEN/f64/lib/mnist_csv2.c" "$GEN/f64/lib/cifar10.c" "$GEN/f64/lib/bmp.c" "$GEN/f64/lib/csv.c" -lm
# lib/mnist_csv.c (mnist_hinge's row reader) apart: its MnistCSV / visualize_digit_data clash with mnist_csv2.c's
$CC $CFLAGS -shared -Wl,-Bsymbolic -o "$OUT/libref_mnist_csv.so" "$GEN/f64/lib/mnist_csv.c" -lm
mk_tree f64_convfix double 1;  build_lib f64_convfix 0
mk_tree f32 float 0;           build_lib f32 1
mk_tree f32_convfix float 1;   build_lib f32_convfix 1

# ---- reference model programs, built exactly as the reference would (its own lib objects) ----
mkdir -p "$OUT/bin"
t="$GEN/f32"
$CC $CFLAGS -o "$OUT/bin/ref_main_f32"           "$t/model/main.c" "$t/lib/matrix.c" "$t/lib/csv.c" "$t/lib/layer.c" -lm -I"$t" 2>/dev/null || \
  ( cd "$t" && mkdir -p top && ln -sf "$REF/main.c" top/main.c && ln -sfn "$t/lib" top/lib && \
    $CC $CFLAGS -o "$OUT/bin/ref_main_f32" top/main.c lib/matrix.c lib/csv.c lib/layer.c -lm )
$CC $CFLAGS -o "$OUT/bin/ref_my_first_model_f32" "$t/model/my_first_model.c" "$t/lib/matrix.c" "$t/lib/csv.c" "$t/lib/layer.c" -lm
$CC $CFLAGS -o "$OUT/bin/ref_mnist_hinge_f32"    "$t/model/mnist_hinge.c" "$t/lib/matrix.c" "$t/lib/csv.c" "$t/lib/mnist_csv.c" -lm
t="$GEN/f64"
$CC $CFLAGS -o "$OUT/bin/ref_mnist_nn_f64"       "$t/model/mnist_nn.c" "$t/lib/matrix.c" "$t/lib/csv.c" "$t/lib/mnist_csv2.c" -lm
# B=512 variant of mnist_nn (SURVEY §8c ref_f64_bN): SGD_BATCH_SIZE is a plain #define
sed 's/^#define SGD_BATCH_SIZE 64/#define SGD_BATCH_SIZE 512/' "$REF/model/mnist_nn.c" > "$t/model/mnist_nn_b512.c"
$CC $CFLAGS -o "$OUT/bin/ref_mnist_nn_f64_b512"  "$t/model/mnist_nn_b512.c" "$t/lib/matrix.c" "$t/lib/csv.c" "$t/lib/mnist_csv2.c" -lm

# ---- the same UNCHANGED model sources relinked against libbla.so (the drop-in demonstration) ----
# lib/{matrix,layer,conv,norm,util}.h come from include/lib (this repo); every other header and the
# host-I/O objects (csv, mnist_csv*, cifar10, bmp) are the reference's own.  See INTEGRATION.md.
BLA_DIR="$HERE/../big-linear-algebra_b200"
if [ -f "$BLA_DIR/libbla.so" ]; then
  t="$GEN/bla"
  mkdir -p "$t/lib" "$t/model"
  for f in "$REF"/lib/*; do ln -s "$f" "$t/lib/$(basename "$f")"; done
  for h in matrix.h layer.h conv.h norm.h util.h; do ln -sf "$HERE/../include/lib/$h" "$t/lib/$h"; done
  for f in "$REF"/model/*.c "$REF"/main.c; do ln -s "$f" "$t/model/$(basename "$f")"; done
  LINK="-L$BLA_DIR -lbla -Wl,-rpath,\$ORIGIN/../../../big-linear-algebra_b200 -lm"
  HOSTFLAGS="-O2 -std=c99 -w"
  $CC $HOSTFLAGS -o "$OUT/bin/bla_main"           "$t/model/main.c" "$t/lib/csv.c" -I"$t" $LINK
  $CC $HOSTFLAGS -o "$OUT/bin/bla_my_first_model" "$t/model/my_first_model.c" "$t/lib/csv.c" $LINK
  $CC $HOSTFLAGS -o "$OUT/bin/bla_mnist_hinge"    "$t/model/mnist_hinge.c" "$t/lib/csv.c" "$t/lib/mnist_csv.c" $LINK
  $CC $HOSTFLAGS -o "$OUT/bin/bla_mnist_nn"       "$t/model/mnist_nn.c" "$t/lib/csv.c" "$t/lib/mnist_csv2.c" $LINK
  sed 's/^#define SGD_BATCH_SIZE 64/#define SGD_BATCH_SIZE 512/' "$REF/model/mnist_nn.c" > "$t/model/mnist_nn_b512.c"
  $CC $HOSTFLAGS -o "$OUT/bin/bla_mnist_nn_b512"  "$t/model/mnist_nn_b512.c" "$t/lib/csv.c" "$t/lib/mnist_csv2.c" $LINK
  $CC $HOSTFLAGS -o "$OUT/bin/bla_cifar_unet"     "$t/model/cifar_unet.c" "$t/lib/csv.c" "$t/lib/cifar10.c" "$t/lib/bmp.c" $LINK
  # the same programs WITHOUT the reference's csv.c: read_csv_contents / write_csv_contents / read_csv_contents_file come from
  # libbla.so's parallel codec (csrc/csv_codec.cu), also underneath the reference's own mnist_csv2.c loader
  $CC $HOSTFLAGS -o "$OUT/bin/bla_main_nocsv"          "$t/model/main.c" -I"$t" $LINK
  $CC $HOSTFLAGS -o "$OUT/bin/bla_mnist_nn_b512_nocsv" "$t/model/mnist_nn_b512.c" "$t/lib/mnist_csv2.c" $LINK
  # ... and with NO reference object or header under lib/ at all: every lib/*.h is this repo's (include/lib), the host I/O
  # (csv, mnist_csv2, cifar10, bmp) comes from libbla.so, mnist_hinge's row reader from libbla_mnist_csv.so
  t2="$GEN/bla_only"
  mkdir -p "$t2/lib" "$t2/model"
  for h in "$HERE"/../include/lib/*.h; do ln -s "$h" "$t2/lib/$(basename "$h")"; done
  for f in "$REF"/model/*.c "$REF"/main.c; do ln -s "$f" "$t2/model/$(basename "$f")"; done
  cp "$t/model/mnist_nn_b512.c" "$t2/model/mnist_nn_b512.c"
  $CC $HOSTFLAGS -o "$OUT/bin/bla_only_main"           "$t2/model/main.c" -I"$t2" $LINK
  $CC $HOSTFLAGS -o "$OUT/bin/bla_only_my_first_model" "$t2/model/my_first_model.c" $LINK
  $CC $HOSTFLAGS -o "$OUT/bin/bla_only_mnist_hinge"    "$t2/model/mnist_hinge.c" -L"$BLA_DIR" -lbla_mnist_csv $LINK
  $CC $HOSTFLAGS -o "$OUT/bin/bla_only_mnist_nn_b512"  "$t2/model/mnist_nn_b512.c" $LINK
  $CC $HOSTFLAGS -o "$OUT/bin/bla_only_cifar_unet"     "$t2/model/cifar_unet.c" $LINK
  # the reference's own U-Net build (double, as shipped) for comparison runs
  $CC $CFLAGS -o "$OUT/bin/ref_cifar_unet_f64" "$GEN/f64/model/cifar_unet.c" "$GEN/f64/lib/matrix.c" "$GEN/f64/lib/csv.c" \
      "$GEN/f64/lib/cifar10.c" "$GEN/f64/lib/bmp.c" "$GEN/f64/lib/conv.c" "$GEN/f64/lib/norm.c" "$GEN/f64/lib/util.c" -lm
  # checkpoint round trip through the reference's own load_parameters / save_parameters (cifar_unet.c:1545-1802; its main() never
  # calls load_parameters, :1902): reads data/cifar_unet under the current directory and writes it back.  For the U-Net checkpoint
  # interoperability test (tests/test_checkpoint_gpu.py).
  printf '%s\n' '#define main ref_cifar_unet_main' '#include "cifar_unet.c"' '#undef main' \
      'int main(void) { ModelParams p; allocate_model_params(&p); load_parameters(&p); save_parameters(&p); return 0; }' > "$GEN/f64/model/unet_ckpt.c"
  $CC $CFLAGS -o "$OUT/bin/ref_unet_ckpt_f64" "$GEN/f64/model/unet_ckpt.c" "$GEN/f64/lib/matrix.c" "$GEN/f64/lib/csv.c" \
      "$GEN/f64/lib/cifar10.c" "$GEN/f64/lib/bmp.c" "$GEN/f64/lib/conv.c" "$GEN/f64/lib/norm.c" "$GEN/f64/lib/util.c" -lm
else
  echo "build_ref: $BLA_DIR/libbla.so not built yet -- skipping the relinked programs" >&2
fi

rm -rf "$GEN"
echo "build_ref: built $(ls "$OUT"/*.so | wc -l) oracle libraries and $(ls "$OUT/bin" | wc -l) reference programs in $OUT"
