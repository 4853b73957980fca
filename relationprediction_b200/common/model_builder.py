"""Name= -> component chain factory (reference: common/model_builder.py:26-184, :273-319).

Only the branches on the accelerated path are built: encoders `gcn_basis` (BasisGcn, or ConcatGcn
when Concatenation=Yes) and `embedding`; decoder `bilinear-diag`.  Unknown names return None exactly
like the reference (:270, :320); ablation flags that select out-of-scope variants raise."""
from ..decoders.bilinear_diag import BilinearDiag
from ..encoders.affine_transform import AffineTransform
from ..encoders.message_gcns.gcn_basis import BasisGcn
from ..encoders.message_gcns.gcn_basis_concat import ConcatGcn
from ..encoders.relation_embedding import RelationEmbedding
from ..extras.graph_representations import Representation


def _flag(settings, key, default="No"):
    return settings[key] if key in settings else default


def build_encoder(encoder_settings, triples):
    name = encoder_settings['Name']
    if name == "embedding":
        input_shape = [int(encoder_settings['EntityCount']), int(encoder_settings['CodeDimension'])]
        embedding = AffineTransform(input_shape, encoder_settings, onehot_input=True, use_bias=False,
                                    use_nonlinearity=False)
        return RelationEmbedding(input_shape, encoder_settings, next_component=embedding)

    if name == "gcn_basis":
        graph = Representation(triples, encoder_settings)
        d_int = int(encoder_settings['InternalEncoderDimension'])
        input_shape = [int(encoder_settings['EntityCount']), d_int]
        internal_shape = [d_int, d_int]
        projection_shape = [d_int, int(encoder_settings['CodeDimension'])]
        relation_shape = [int(encoder_settings['EntityCount']), int(encoder_settings['CodeDimension'])]
        layers = int(encoder_settings['NumberOfLayers'])

        if _flag(encoder_settings, 'UseInputTransform') != "Yes":
            raise NotImplementedError("UseInputTransform=No / RandomInput / PartiallyRandomInput variants are "
                                      "outside the accelerated path (SURVEY.md 2.1 #6)")
        encoding = AffineTransform(input_shape, encoder_settings, next_component=graph, onehot_input=True,
                                   use_bias=True, use_nonlinearity=True)
        encoding = apply_basis_gcn(encoder_settings, encoding, internal_shape, layers)
        if _flag(encoder_settings, 'UseOutputTransform') == "Yes":
            encoding = AffineTransform(projection_shape, encoder_settings, next_component=encoding,
                                       onehot_input=False, use_nonlinearity=False, use_bias=True)
        return RelationEmbedding(relation_shape, encoder_settings, next_component=encoding)
    return None


def apply_basis_gcn(encoder_settings, encoding, internal_shape, layers):
    for flag in ('AddDiagonal', 'DiagonalCoefficients', 'StoreEdgeData'):
        if _flag(encoder_settings, flag) == "Yes":
            raise NotImplementedError("%s=Yes selects an ablation variant outside the accelerated path" % flag)
    if _flag(encoder_settings, 'SkipConnections', 'None') not in ('None', 'Residual'):
        # 'Residual' is a no-op in the reference (model_builder.py:302-307 overwrites it); 'Highway' is not
        raise NotImplementedError("SkipConnections=Highway is outside the accelerated path")
    model = ConcatGcn if _flag(encoder_settings, 'Concatenation') == "Yes" else BasisGcn
    for layer in range(layers):
        use_nonlinearity = layer < layers - 1  # the last layer is linear (model_builder.py:275)
        encoding = model(internal_shape, encoder_settings, next_component=encoding, onehot_input=False,
                         use_nonlinearity=use_nonlinearity)
    return encoding


def build_decoder(encoder, decoder_settings):
    if decoder_settings['Name'] == "bilinear-diag":
        return BilinearDiag(encoder, decoder_settings)
    return None
